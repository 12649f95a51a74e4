// comm.cuh -- the two exchange steps of the path across GPUs: the depth a shard's reads leave beyond its range end
// goes to the next rank (prefix-sum carry across a shard boundary), and the genome-wide `total` histograms / stats are all-reduced
// (sum_into mosdepth.nim:761-764, depth_stat `+` depthstat.nim:53-57).  NCCL over NVLink; nothing in the single-process reference
// corresponds to this file.
//
// NCCL is bound at run time (dlopen "libnccl.so.2", the soname both the system library and the one PyTorch ships carry): the
// library has no link-time dependency on it, loads on a box without NCCL, and inside a Python process that imported torch first it
// shares torch's copy.  Only the declarations of <nccl.h> are used.
#pragma once
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

#include "common.cuh"

namespace msd {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    char why[256] = "";
};

inline NcclApi* nccl_api() {
    static NcclApi api; static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("MSD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { if (!n || !*n) continue; api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) { snprintf(api.why, sizeof api.why, "cannot load libnccl.so.2: %s", dlerror()); return; }
        bool ok = true;
        auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) { ok = false; snprintf(api.why, sizeof api.why, "libnccl has no %s", s); } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return &api;      // handle == nullptr: unavailable, `why` says why
}

// ---- spill message: [0] = number of entries n, then n x {tid, pos, cells}, then the cells of all entries back to back -----------
// One fixed-size message per neighbour pair (the receiver cannot know the size in advance, and a second round trip would cost
// more than the bytes: 4 MB cross NVLink in ~10 us).
constexpr uint32_t kSpillHeaderWords = 1024;            // up to 341 entries per rank
constexpr uint32_t kSpillMaxEntries = (kSpillHeaderWords - 1) / 3;

// sender: copy the cells of every entry out of the depth arena behind the header (the header itself comes from the host)
__global__ void spill_pack_kernel(const int32_t* __restrict__ arena, const unsigned long long* __restrict__ depth_off, uint32_t* __restrict__ msg) {
    const uint32_t n = msg[0];
    uint32_t base = kSpillHeaderWords;
    for (uint32_t e = 0; e < n; e++) {
        const uint32_t tid = msg[1 + 3 * e], pos = msg[2 + 3 * e], cells = msg[3 + 3 * e];
        const int32_t* __restrict__ src = arena + depth_off[tid] + pos;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) msg[base + i] = (uint32_t)src[i];
        base += cells;
    }
}

// receiver: add every entry at the start of the range this context owns of that contig.  An entry that does not meet an owned range
// starting exactly there, or that is longer than the range, raises *error (1: no such range, 2: longer than the receiving shard).
__global__ void spill_unpack_kernel(int32_t* __restrict__ arena, const unsigned long long* __restrict__ depth_off, const uint32_t* __restrict__ tlen,
                                    const uint32_t* __restrict__ range_lo, const uint32_t* __restrict__ range_hi, int n_targets,
                                    const uint32_t* __restrict__ msg, uint32_t msg_words, uint32_t* __restrict__ error) {
    const uint32_t n = msg[0];
    if (n > kSpillMaxEntries) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(error, 3u); return; }
    uint32_t base = kSpillHeaderWords;
    for (uint32_t e = 0; e < n; e++) {
        const uint32_t tid = msg[1 + 3 * e], pos = msg[2 + 3 * e], cells = msg[3 + 3 * e];
        if (tid >= (uint32_t)n_targets || depth_off[tid] == ~0ull || range_lo[tid] != pos || pos == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(error, 1u); return; }
        const uint32_t hi = range_hi[tid] < tlen[tid] ? range_hi[tid] : tlen[tid];
        if ((unsigned long long)pos + cells > (unsigned long long)hi + (hi == tlen[tid] ? 1u : 0u) || (unsigned long long)base + cells > msg_words) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(error, 2u); return; }
        int32_t* __restrict__ dst = arena + depth_off[tid] + pos;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) dst[i] += (int32_t)msg[base + i];
        base += cells;
    }
}

// the eight stats {cum_length, cum_depth, min, max} x {global, region} -> [sum x 4 | min x 2 | max x 2] and back
__global__ void stats_pack_kernel(const long long* __restrict__ s, long long* __restrict__ p) { p[0] = s[0]; p[1] = s[1]; p[2] = s[4]; p[3] = s[5]; p[4] = s[2]; p[5] = s[6]; p[6] = s[3]; p[7] = s[7]; }
__global__ void stats_unpack_kernel(long long* __restrict__ s, const long long* __restrict__ p) { s[0] = p[0]; s[1] = p[1]; s[4] = p[2]; s[5] = p[3]; s[2] = p[4]; s[6] = p[5]; s[3] = p[6]; s[7] = p[7]; }

}  // namespace msd
