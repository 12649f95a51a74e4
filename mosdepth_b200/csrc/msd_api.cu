// msd_api.cu -- C ABI of libmosdepth_b200.so (include/mosdepth_b200.h): context, chunk pipeline, query calls.
// Host code here only plans work (BSIZE chase, chunk descriptors) and moves bytes; every byte of decode, filter,
// accumulate and reduce work runs in the sm_100a kernels of the *.cuh files.  There is no CPU fallback.
#include "../../include/mosdepth_b200.h"

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "common.cuh"
#include "frame.cuh"
#include "inflate.cuh"
#include "records.cuh"
#include "regions.cuh"
#include "runs.cuh"
#include "scan.cuh"
#include "bedgz_host.cuh"
#include "comm.cuh"
#include "ingest.hpp"

using namespace msd;

namespace {

thread_local char g_create_error[512] = "";

struct Span { uint64_t beg, end; };

// Chunks in flight.  One decode wave costs the same ~7 ms for 6,000 blocks as for 42,624 (the chain per block is serial), so
// the pipeline wants several modest chunks whose decode launches share the SMs, not two big ones run back to back.
constexpr int kSlots = 8;       // most chunk slots a context can have; ctx->n_slots (MSD_SLOTS = 2, 4 or 8; default 4) are in use
constexpr int kDescSets = 16;   // block descriptors: deeper than the slots, so that the copy stream never waits for a chunk's framing
                                // (a multiple of n_slots: chunk i and chunk i - kDescSets share a slot and so an inflate stream)

struct StageSet {              // pinned host staging for one in-flight chunk
    BlockDesc* h_desc = nullptr; uint32_t* h_rel = nullptr;
    uint8_t* h_bounce = nullptr;   // pinned copy of the compressed bytes when the source is pageable
    cudaEvent_t free_ev = nullptr; // recorded when the device no longer needs this set
    bool used = false;
};

struct ContigResult {
    bool have = false;
    msd_depth_stat stat{};
    std::vector<int64_t> dist;
};

}  // namespace

struct msd_ctx {
    int device = 0;
    char err[512] = "";
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_inflate[kSlots] = {}; int n_slots = 4;
    // input
    const uint8_t* bytes = nullptr; uint64_t n_bytes = 0; bool mapped = false; bool src_pinned = false;
    int n_targets = 0; std::vector<uint32_t> tlen; uint64_t first_voff = 0;
    std::vector<Span> spans; std::vector<uint8_t> want;
    // staged compressed bytes
    uint8_t* d_staged = nullptr; uint64_t staged_beg = 0, staged_end = 0;
    // filters
    FilterParams fp{}; std::string rg_blob; char* d_rg = nullptr; bool filters_set = false;
    // depth arena
    int32_t* d_arena = nullptr; uint64_t arena_cells = 0; std::vector<uint64_t> depth_off;
    uint64_t* d_depth_off = nullptr; uint32_t* d_tlen = nullptr;
    // intra-contig sharding (msd_set_contig_range)
    std::vector<uint32_t> range_lo, range_hi, max_end; std::vector<uint8_t> ranged, pending; bool any_ranged = false;
    uint32_t *d_range_lo = nullptr, *d_range_hi = nullptr, *d_max_end = nullptr; unsigned long long* d_n_beyond = nullptr;
    uint32_t* d_min_need = nullptr; std::vector<uint32_t> min_need, covered_from;   // soundness of a shard's span start (default mode, msd_set_contig_cover)
    unsigned long long* d_nprefilter = nullptr; std::vector<long long> n_prefilter;
    std::vector<ContigResult> results; bool ran = false;
    // chunk buffers
    uint64_t comp_cap = 0, infl_cap = 0; uint32_t carry_cap = 0, max_blocks = 0, rec_cap = 0;
    uint8_t* d_comp[kSlots] = {}; uint8_t* d_infl[kSlots] = {};
    BlockDesc* d_desc[kDescSets] = {}; uint32_t* d_rel[kDescSets] = {};
    uint32_t *d_guess = nullptr, *d_seg = nullptr, *d_bstart = nullptr, *d_land = nullptr, *d_count = nullptr, *d_first = nullptr, *d_cnt = nullptr, *d_recbase = nullptr, *d_recoff = nullptr;
    uint32_t *d_status = nullptr, *d_ticket = nullptr, *d_inflate_err = nullptr;
    uint32_t* d_tokens[kSlots] = {}; uint32_t* d_ntok[kSlots] = {};   // match tokens between K1a and K1b
    uint32_t* d_crc[kSlots] = {}; bool check_crc = true;
    unsigned long long* d_ring[kSlots] = {};                          // fused K1 (MSD_FUSED=1): blocks announced by decode lanes, taken by resolve warps              // CRC32 fields of the gzip trailers (K1c)
    ChunkState* d_cs = nullptr; RunCounters* d_rc = nullptr;
    StageSet stage[kDescSets];   // h_desc / h_rel per descriptor set; h_bounce in the first kSlots sets only
    cudaEvent_t desc_free[kDescSets] = {}, crc_done[kDescSets] = {}, bounce_free[kSlots] = {}; bool bounce_used[kSlots] = {};
    cudaEvent_t comp_free[kSlots] = {}, h2d_done[kSlots] = {}, infl_done[kSlots] = {}, infl_free[kSlots] = {};
    // mate join (default mode only)
    MateChunk mc{}; MateTable mt{}; LiveList live[2]{}; bool mate_ready = false; uint32_t mate_slots = 0;
    // scan + queries
    unsigned long long* d_tile_state = nullptr; uint64_t tile_state_cap = 0; uint32_t* d_scan_ticket = nullptr;
    unsigned long long* d_hist = nullptr; ContigAccum* d_acc = nullptr; RegionAccum* d_racc = nullptr;
    // per-contig results of msd_run gathered without a host round trip per contig
    // K6: every segment (contig or shard) of a run in one launch; results per slot, gathered with one asynchronous copy
    ScanSeg *d_segs = nullptr, *h_segs = nullptr; ContigAccum *d_acc_all = nullptr, *h_acc_all = nullptr;
    unsigned long long *d_hist_all = nullptr, *h_hist_all = nullptr; int segs_cap = 0;
    unsigned long long* d_dist512 = nullptr; unsigned long long* d_rdist = nullptr;
    // totals for multi-GPU reduction
    long long* d_tot_global = nullptr; long long* d_tot_region = nullptr; long long* d_tot_stats = nullptr;
    // run buffers (library-owned pinned host memory)
    int32_t *h_run_start = nullptr, *h_run_stop = nullptr, *h_run_val = nullptr; uint64_t h_run_cap = 0;
    int32_t *d_run_start = nullptr, *d_run_stop = nullptr, *d_run_val = nullptr; uint64_t d_run_cap = 0;
    uint32_t* d_tile_count = nullptr; unsigned long long* d_tile_base = nullptr; uint64_t run_tile_cap = 0; unsigned long long* d_run_total = nullptr;
    int32_t *d_keep_start = nullptr, *d_keep_stop = nullptr, *d_keep_val = nullptr; uint64_t d_keep_cap = 0;      // K9's filter compacts into these, then the two sets swap
    // K10 (msd_per_base_bgzf): grow-only buffers (per line, per member + text, output) and the pinned copy of the output
    BzBuffers bz;
    // generic scratch
    void* d_scratch = nullptr; uint64_t scratch_cap = 0;
    // msd_prefetch_windows: every contig's windows computed by one launch at the end of msd_run, results cached on the host
    uint32_t pf_window = 0; bool pf_valid = false; uint64_t pf_cap_windows = 0; int pf_cap_contigs = 0;
    double *d_pf_values = nullptr, *h_pf_values = nullptr; RegionAccum *d_pf_acc = nullptr, *h_pf_acc = nullptr;
    unsigned long long *d_pf_dist = nullptr, *h_pf_dist = nullptr; WinContig *d_pf_table = nullptr, *h_pf_table = nullptr;
    std::vector<int> pf_slot;   // tid -> row of the tables (-1: not prefetched)
    // multi-GPU (comm.cuh): one communicator per context; fixed-size spill messages to / from the neighbouring ranks
    ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1;
    uint32_t *d_spill_send = nullptr, *d_spill_recv = nullptr, *h_spill_hdr = nullptr, *d_comm_err = nullptr; uint32_t spill_words = 0; long long* d_stats_pack = nullptr;
    IngestPump* pump = nullptr;   // pageable input: pinned-ring copy thread (ingest.hpp), created by the first msd_run that needs it
    msd_run_info info{};
    uint64_t launches = 0; cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
};

namespace {

int ctx_fail(msd_ctx* ctx, int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    if (ctx) vsnprintf(ctx->err, sizeof ctx->err, fmt, ap); else vsnprintf(g_create_error, sizeof g_create_error, fmt, ap);
    va_end(ap);
    return code;
}

template <class T> int dev_alloc(msd_ctx* ctx, T** p, uint64_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "cudaMalloc(%llu bytes): %s", (unsigned long long)(n * sizeof(T)), cudaGetErrorString(e));
    return 0;
}
template <class T> void dev_free(T** p) { if (*p) { cudaFree(*p); *p = nullptr; } }

cudaEvent_t next_event(msd_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
    return ctx->ev_pool[ctx->ev_used++];
}

uint64_t env_u64(const char* name, uint64_t dflt) { const char* v = getenv(name); return (v && *v) ? strtoull(v, nullptr, 10) : dflt; }

void release_input(msd_ctx* ctx) {
    if (ctx->pump) { delete ctx->pump; ctx->pump = nullptr; }
    if (ctx->mapped && ctx->bytes) munmap((void*)ctx->bytes, ctx->n_bytes);
    ctx->bytes = nullptr; ctx->n_bytes = 0; ctx->mapped = false; ctx->src_pinned = false;
    dev_free(&ctx->d_staged); ctx->staged_beg = ctx->staged_end = 0;
    ctx->ran = false;
}

int grid_for(uint64_t n, int threads, int max_ctas) { uint64_t g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (uint64_t)max_ctas) g = max_ctas; return (int)g; }

// parse one BGZF header at `pos`; returns block length or 0 on error
uint32_t bgzf_header(const uint8_t* f, uint64_t n, uint64_t pos, uint32_t* xlen_out, uint32_t* isize_out) {
    if (pos + 18 > n) return 0;
    const uint8_t* h = f + pos;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    uint32_t xlen = h[10] | (h[11] << 8), q = 12, bsize = 0; bool have = false;
    if (pos + 12 + xlen > n) return 0;
    while (q + 4 <= 12 + xlen) {
        uint32_t slen = h[q + 2] | (h[q + 3] << 8);
        if (h[q] == 66 && h[q + 1] == 67 && slen == 2 && q + 6 <= 12 + xlen) { bsize = h[q + 4] | (h[q + 5] << 8); have = true; }
        q += 4 + slen;
    }
    if (!have) return 0;
    uint32_t blen = bsize + 1;
    if (blen < xlen + 20 || pos + blen > n) return 0;
    memcpy(isize_out, f + pos + blen - 4, 4);
    *xlen_out = xlen;
    return blen;
}

int setup_common(msd_ctx* ctx, int n_targets, const uint32_t* target_len, uint64_t first_voff) {
    if (n_targets < 0 || (n_targets > 0 && !target_len)) return ctx_fail(ctx, MSD_EINVAL, "bad target list");
    ctx->n_targets = n_targets; ctx->tlen.assign(target_len, target_len + n_targets); ctx->first_voff = first_voff;
    ctx->spans.clear(); ctx->want.assign(n_targets, 1);
    ctx->results.assign(n_targets, ContigResult()); ctx->n_prefilter.assign(n_targets, 0);
    ctx->range_lo.assign(n_targets, 0); ctx->range_hi.assign(n_targets, 0xffffffffu); ctx->max_end.assign(n_targets, 0);
    ctx->ranged.assign(n_targets, 0); ctx->pending.assign(n_targets, 0); ctx->any_ranged = false;
    ctx->covered_from.assign(n_targets, 0xffffffffu); ctx->min_need.assign(n_targets, 0xffffffffu);
    cudaPointerAttributes pa; memset(&pa, 0, sizeof pa);
    ctx->src_pinned = (cudaPointerGetAttributes(&pa, ctx->bytes) == cudaSuccess && pa.type == cudaMemoryTypeHost);
    cudaGetLastError();
    // chunk geometry
    ctx->comp_cap = std::min<uint64_t>(env_u64("MSD_CHUNK_COMP_MB", 512) << 20, ((ctx->n_bytes + 65536 + 255) / 256) * 256);
    ctx->infl_cap = std::min<uint64_t>(env_u64("MSD_CHUNK_INFL_MB", 1600) << 20, std::max<uint64_t>(8ull << 20, ctx->n_bytes * 16));
    ctx->carry_cap = (uint32_t)std::min<uint64_t>(env_u64("MSD_CARRY_MB", 64) << 20, ctx->infl_cap);
    ctx->max_blocks = (uint32_t)std::min<uint64_t>(env_u64("MSD_CHUNK_BLOCKS", 24000), ctx->n_bytes / 28 + 2);
    ctx->rec_cap = (uint32_t)((ctx->infl_cap + ctx->carry_cap) / 36 + 2);   // the smallest record framing accepts: block_size 32 = 36 bytes
    return 0;
}

// chunk buffers, staging and their events depend on the input size: dropped when a new input is opened and on destroy
void free_chunk_buffers(msd_ctx* ctx) {
    for (int k = 0; k < kDescSets; k++) {
        dev_free(&ctx->d_desc[k]); dev_free(&ctx->d_rel[k]);
        if (ctx->stage[k].h_desc) { cudaFreeHost(ctx->stage[k].h_desc); ctx->stage[k].h_desc = nullptr; }
        if (ctx->stage[k].h_rel) { cudaFreeHost(ctx->stage[k].h_rel); ctx->stage[k].h_rel = nullptr; }
        if (ctx->stage[k].h_bounce) { cudaFreeHost(ctx->stage[k].h_bounce); ctx->stage[k].h_bounce = nullptr; }
        cudaEvent_t* evs[] = {&ctx->stage[k].free_ev, &ctx->desc_free[k], &ctx->crc_done[k]};
        for (cudaEvent_t* e : evs) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
        ctx->stage[k].used = false;
    }
    for (int k = 0; k < ctx->n_slots; k++) {
        dev_free(&ctx->d_comp[k]); dev_free(&ctx->d_infl[k]); dev_free(&ctx->d_tokens[k]); dev_free(&ctx->d_ntok[k]); dev_free(&ctx->d_crc[k]); dev_free(&ctx->d_ring[k]);
        cudaEvent_t* evs[] = {&ctx->comp_free[k], &ctx->h2d_done[k], &ctx->infl_done[k], &ctx->infl_free[k], &ctx->bounce_free[k]};
        for (cudaEvent_t* e : evs) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
        ctx->bounce_used[k] = false;
    }
}

int ensure_chunk_buffers(msd_ctx* ctx) {
    if (ctx->d_infl[0]) return 0;
    int rc;
    const uint64_t infl_bytes = (uint64_t)ctx->carry_cap + ctx->infl_cap + 256;
    for (int k = 0; k < ctx->n_slots; k++) {
        if ((rc = dev_alloc(ctx, &ctx->d_comp[k], ctx->comp_cap + 256))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_infl[k], infl_bytes))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_tokens[k], (uint64_t)ctx->max_blocks * kTokenCap))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_ntok[k], ctx->max_blocks))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_crc[k], ctx->max_blocks))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_ring[k], ctx->max_blocks))) return rc;
        MSD_CUDA_OK(cudaMemset(ctx->d_infl[k], 0, infl_bytes));
        cudaEvent_t* evs[] = {&ctx->comp_free[k], &ctx->h2d_done[k], &ctx->infl_done[k], &ctx->infl_free[k], &ctx->bounce_free[k]};
        for (cudaEvent_t* e : evs) MSD_CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    for (int k = 0; k < kDescSets; k++) {
        if ((rc = dev_alloc(ctx, &ctx->d_desc[k], ctx->max_blocks))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_rel[k], ctx->max_blocks + 1))) return rc;
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->stage[k].h_desc, sizeof(BlockDesc) * ctx->max_blocks));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->stage[k].h_rel, sizeof(uint32_t) * (ctx->max_blocks + 1)));
        cudaEvent_t* evs[] = {&ctx->stage[k].free_ev, &ctx->desc_free[k], &ctx->crc_done[k]};
        for (cudaEvent_t* e : evs) MSD_CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    const uint32_t nb = ctx->max_blocks + 2;
    if ((rc = dev_alloc(ctx, &ctx->d_bstart, nb + 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_land, nb))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_guess, nb + 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_seg, nb + 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_count, nb))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_first, nb))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_cnt, nb))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_recbase, nb))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_recoff, ctx->rec_cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_status, (uint64_t)ctx->n_slots * ctx->max_blocks))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_ticket, 4 * kSlots))) return rc;
    ctx->check_crc = !getenv("MSD_NO_CRC");
    if ((rc = dev_alloc(ctx, &ctx->d_inflate_err, 4))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_cs, 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_rc, 1))) return rc;
    return 0;
}

int ensure_mate(msd_ctx* ctx) {
    if (ctx->mate_ready) return 0;
    int rc;
    const uint64_t R = ctx->rec_cap;
    if ((rc = dev_alloc(ctx, &ctx->mc.cls, R))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mc.qhash, R))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mc.endpos, R))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mc.next, R))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mc.slot, R))) return rc;
    const uint32_t live_cap = (uint32_t)std::min<uint64_t>(env_u64("MSD_LIVE_CAP", 8u << 20), std::max<uint64_t>(R * 4, 1024));
    uint32_t slots = 1024; while (slots < 2ull * (R + live_cap) && slots < (1u << 26)) slots <<= 1;
    slots = (uint32_t)std::min<uint64_t>(slots, env_u64("MSD_MATE_SLOTS", slots));
    ctx->mate_slots = slots;
    if ((rc = dev_alloc(ctx, &ctx->mt.word, slots))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mt.head, slots))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->mt.live, slots))) return rc;
    ctx->mt.mask = slots - 1;
    const uint32_t arena_cap = (uint32_t)std::min<uint64_t>(env_u64("MSD_LIVE_ARENA_MB", 512) << 20, std::max<uint64_t>((uint64_t)live_cap * 64, 1 << 20));
    for (int k = 0; k < 2; k++) {
        if ((rc = dev_alloc(ctx, &ctx->live[k].e, live_cap))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->live[k].arena, arena_cap))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->live[k].n, 1))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->live[k].arena_used, 1))) return rc;
        ctx->live[k].cap = live_cap; ctx->live[k].arena_cap = arena_cap;
    }
    ctx->mate_ready = true;
    return 0;
}

int ensure_arena(msd_ctx* ctx) {
    // every selected contig gets tlen+1 cells (mosdepth.nim:274), padded to 128 B so 128-bit accesses stay aligned
    std::vector<uint64_t> off(ctx->n_targets, ~0ull);
    uint64_t cells = 0;
    for (int t = 0; t < ctx->n_targets; t++) if (ctx->want[t]) { off[t] = cells; cells += (((uint64_t)ctx->tlen[t] + 1 + 31) / 32) * 32; }
    int rc;
    if (cells > ctx->arena_cells || !ctx->d_arena) { if ((rc = dev_alloc(ctx, &ctx->d_arena, cells + 32))) return rc; ctx->arena_cells = cells; }
    ctx->depth_off = off;
    if (!ctx->d_depth_off) {
        if ((rc = dev_alloc(ctx, &ctx->d_depth_off, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_tlen, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_nprefilter, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_range_lo, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_range_hi, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_max_end, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_n_beyond, (uint64_t)ctx->n_targets))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_min_need, (uint64_t)ctx->n_targets))) return rc;
    }
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_range_lo, ctx->range_lo.data(), sizeof(uint32_t) * ctx->n_targets, cudaMemcpyHostToDevice, ctx->s_compute));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_range_hi, ctx->range_hi.data(), sizeof(uint32_t) * ctx->n_targets, cudaMemcpyHostToDevice, ctx->s_compute));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_depth_off, off.data(), sizeof(uint64_t) * ctx->n_targets, cudaMemcpyHostToDevice, ctx->s_compute));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_tlen, ctx->tlen.data(), sizeof(uint32_t) * ctx->n_targets, cudaMemcpyHostToDevice, ctx->s_compute));
    MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));   // `off` is a stack vector
    return 0;
}

int ensure_query_buffers(msd_ctx* ctx) {
    if (ctx->d_hist) return 0;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->d_hist, (uint64_t)kDistBins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_rdist, (uint64_t)kDistBins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_dist512, (uint64_t)kWindowDistBins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_acc, 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_racc, 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_scan_ticket, 4))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_tot_global, (uint64_t)kDistBins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_tot_region, (uint64_t)kDistBins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_tot_stats, 8))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_run_total, 1))) return rc;
    return msd_totals_reset(ctx);
}

int ensure_scratch(msd_ctx* ctx, uint64_t bytes) {
    if (bytes <= ctx->scratch_cap) return 0;
    if (ctx->d_scratch) { cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; }
    cudaError_t e = cudaMalloc(&ctx->d_scratch, bytes);
    if (e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "cudaMalloc scratch %llu: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    ctx->scratch_cap = bytes;
    return 0;
}

// totals accumulation kernels (sum_into, mosdepth.nim:491-499, and depth_stat `+`, depthstat.nim:53-57)
__global__ void add_hist_kernel(long long* __restrict__ to, const unsigned long long* __restrict__ from, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { unsigned long long v = from[i]; if (v) to[i] += (long long)v; }
}
// sum_into + depth_stat `+` for the window rows of every prefetched contig (one CTA per contig)
__global__ void window_totals_kernel(const RegionAccum* __restrict__ acc, const unsigned long long* __restrict__ dist512, int n_contigs, long long* __restrict__ tot_region, long long* __restrict__ tot_stats4) {
    const int c = blockIdx.x;
    if (c >= n_contigs) return;
    for (int i = threadIdx.x; i < kWindowDistBins; i += blockDim.x) { const unsigned long long v = dist512[(size_t)c * kWindowDistBins + i]; if (v) atomicAdd((unsigned long long*)&tot_region[i], v); }
    if (threadIdx.x == 0) {
        const RegionAccum a = acc[c];
        atomicAdd((unsigned long long*)&tot_stats4[0], a.cum_length); atomicAdd((unsigned long long*)&tot_stats4[1], a.cum_depth);
        atomicMin(&tot_stats4[2], (long long)a.min_depth); atomicMax(&tot_stats4[3], (long long)a.max_depth);
    }
}
__global__ void init_acc_all_kernel(ContigAccum* a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i].cum_depth = 0; a[i].min_depth = 0xffffffffu; a[i].max_depth = 0; a[i].neg_seen = 0; a[i].pad = 0; }
}

__global__ void add_stats_kernel(long long* __restrict__ tot, long long cum_length, unsigned long long cum_depth, unsigned int mn, unsigned int mx) {
    tot[0] += cum_length; tot[1] += (long long)cum_depth;
    if ((long long)mn < tot[2]) tot[2] = mn;
    if ((long long)mx > tot[3]) tot[3] = mx;
}
__global__ void init_stats_kernel(long long* __restrict__ tot) { tot[0] = 0; tot[1] = 0; tot[2] = 0xffffffffll; tot[3] = 0; tot[4] = 0; tot[5] = 0; tot[6] = 0xffffffffll; tot[7] = 0; }
__global__ void init_accum_kernel(ContigAccum* a, RegionAccum* r) {
    if (a) { a->cum_depth = 0; a->min_depth = 0xffffffffu; a->max_depth = 0; a->neg_seen = 0; a->pad = 0; }
    if (r) { r->cum_length = 0; r->cum_depth = 0; r->min_depth = 0xffffffffu; r->max_depth = 0; r->neg_seen = 0; r->pad = 0; }
}

struct StageTimes { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> h2d, inflate, frame, records, k1; };   // k1: end of K1a, end of K1b

double sum_ms(const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v) {
    double s = 0; for (auto& p : v) { float ms = 0; if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) s += ms; } return s;
}

// K6 for a batch of segments: prefix sum (scan = true) and / or the whole-segment stats + dist, one launch; the `total`
// accumulators are updated on the device.  Nothing is read back here: collect_segments() does that after the caller's sync.
constexpr int kMaxSegsPerLaunch = 4096;

int ensure_segments(msd_ctx* ctx, int n_segs, uint64_t n_tiles) {
    int rc;
    if (n_segs > ctx->segs_cap) {
        const int cap = std::max(n_segs, 64);
        dev_free(&ctx->d_segs); dev_free(&ctx->d_acc_all); dev_free(&ctx->d_hist_all);
        if (ctx->h_segs) { cudaFreeHost(ctx->h_segs); ctx->h_segs = nullptr; }
        if (ctx->h_acc_all) { cudaFreeHost(ctx->h_acc_all); ctx->h_acc_all = nullptr; }
        if (ctx->h_hist_all) { cudaFreeHost(ctx->h_hist_all); ctx->h_hist_all = nullptr; }
        ctx->segs_cap = 0;
        if ((rc = dev_alloc(ctx, &ctx->d_segs, (uint64_t)cap)) || (rc = dev_alloc(ctx, &ctx->d_acc_all, (uint64_t)cap)) || (rc = dev_alloc(ctx, &ctx->d_hist_all, (uint64_t)cap * kHistKeep))) return rc;
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_segs, sizeof(ScanSeg) * cap));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_acc_all, sizeof(ContigAccum) * cap));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_hist_all, 8ull * kHistKeep * cap));
        ctx->segs_cap = cap;
    }
    if (n_tiles > ctx->tile_state_cap) { if ((rc = dev_alloc(ctx, &ctx->d_tile_state, n_tiles))) return rc; ctx->tile_state_cap = n_tiles; }
    return 0;
}

// segs[i].first_tile / .slot are filled here (slot = i)
int run_segments(msd_ctx* ctx, int32_t* base, std::vector<ScanSeg>& segs, bool scan) {
    const int n = (int)segs.size();
    if (n == 0) return 0;
    if (n > kMaxSegsPerLaunch) return ctx_fail(ctx, MSD_EINVAL, "run_segments: batch too large");
    // The scan launch: 32-KB tiles (256 threads x 8 int4), three CTAs per SM, L2 prefetch of the tile one grid further on (r2n: 0.68 ms for 370 M cells
    // against 0.85 ms with 16-KB tiles and four CTAs per SM, MSD_SCAN=0, kept as the A/B hook; the stats-only pass always uses 16-KB tiles).
    const uint64_t variant = scan ? env_u64("MSD_SCAN", 6) : 0;
    const uint64_t tile_cells = variant == 6 ? (uint64_t)scan_tile_cells(8) : (uint64_t)kScanTile;
    uint64_t n_tiles = 0; bool any_stats = false;
    for (int i = 0; i < n; i++) { segs[i].first_tile = (uint32_t)n_tiles; segs[i].slot = (uint32_t)i; n_tiles += ((uint64_t)segs[i].n_scan + tile_cells - 1) / tile_cells; any_stats |= segs[i].n_stat != 0; }
    if (n_tiles >= 0xffffffffull) return ctx_fail(ctx, MSD_EINVAL, "run_segments: too many tiles");
    int rc = ensure_segments(ctx, n, n_tiles); if (rc) return rc;
    cudaStream_t sc = ctx->s_compute;
    memcpy(ctx->h_segs, segs.data(), sizeof(ScanSeg) * n);
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_segs, ctx->h_segs, sizeof(ScanSeg) * n, cudaMemcpyHostToDevice, sc));
    if (any_stats) {
        init_acc_all_kernel<<<(n + 255) / 256, 256, 0, sc>>>(ctx->d_acc_all, n);
        MSD_CUDA_OK(cudaMemsetAsync(ctx->d_hist_all, 0, 8ull * kHistKeep * n, sc));
    }
    if (n_tiles) {
        const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)kSMs * 8);
        if (scan) {
            MSD_CUDA_OK(cudaMemsetAsync(ctx->d_tile_state, 0, n_tiles * 8, sc));
            MSD_CUDA_OK(cudaMemsetAsync(ctx->d_scan_ticket, 0, 4, sc));
#define MSD_SCAN_LAUNCH(G, V, A, CTAS) scan_segments_kernel<true, G, V, A><<<(int)std::min<uint64_t>(n_tiles, (uint64_t)kSMs * (CTAS)), kScanThreads, 0, sc>>>( \
    base, ctx->d_segs, n, (uint32_t)n_tiles, ctx->d_tile_state, ctx->d_scan_ticket, ctx->d_acc_all, ctx->d_hist_all)
            if (variant == 6) MSD_SCAN_LAUNCH(1, 8, true, 3);
            else MSD_SCAN_LAUNCH(1, 4, false, 4);
#undef MSD_SCAN_LAUNCH
        } else
            scan_segments_kernel<false><<<grid, kScanThreads, 0, sc>>>(base, ctx->d_segs, n, (uint32_t)n_tiles, ctx->d_tile_state, ctx->d_scan_ticket, ctx->d_acc_all, ctx->d_hist_all);
        MSD_CUDA_OK(cudaGetLastError());
        ctx->info.n_launches++;
    }
    if (any_stats) {
        totals_add_kernel<<<kTotalsParts * n, 256, 0, sc>>>(ctx->d_segs, n, ctx->d_acc_all, ctx->d_hist_all, ctx->d_tot_global, ctx->d_tot_stats);
        MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_acc_all, ctx->d_acc_all, sizeof(ContigAccum) * n, cudaMemcpyDeviceToHost, sc));
        MSD_CUDA_OK(cudaGetLastError());
        ctx->info.n_launches += 2;
    }
    return 0;
}

// after the stream has been synchronised: the per-segment results of the last run_segments() -> ctx->results[tid_of[i]]
int collect_segments(msd_ctx* ctx, const int32_t* base, const std::vector<ScanSeg>& segs, const std::vector<int>& tid_of) {
    const int n = (int)segs.size();
    cudaStream_t sc = ctx->s_compute;
    uint32_t width = 0;
    for (int i = 0; i < n; i++) if (segs[i].n_stat) { const uint32_t mx = ctx->h_acc_all[i].max_depth; width = std::max<uint32_t>(width, mx >= (uint32_t)kHistKeep ? 1u : mx + 1); }
    if (width) {
        MSD_CUDA_OK(cudaMemcpy2DAsync(ctx->h_hist_all, 8ull * kHistKeep, ctx->d_hist_all, 8ull * kHistKeep, 8ull * width, (size_t)n, cudaMemcpyDeviceToHost, sc));
        MSD_CUDA_OK(cudaStreamSynchronize(sc));
    }
    for (int i = 0; i < n; i++) {
        if (!segs[i].n_stat) continue;
        const ContigAccum acc = ctx->h_acc_all[i];
        ContigResult& R = ctx->results[tid_of[i]];
        R.have = true;
        R.stat.cum_length = segs[i].n_stat; R.stat.cum_depth = acc.cum_depth; R.stat.min_depth = acc.min_depth; R.stat.max_depth = acc.max_depth;
        // distribution: bins up to the maximum depth seen (signed view: negative depths are skipped by inc())
        const uint32_t mx = acc.max_depth;
        const int64_t top = (mx > 0x7fffffffu) ? (int64_t)kMaxCoverage : (int64_t)std::min<uint32_t>(mx, kMaxCoverage);
        R.dist.assign((size_t)top + 1, 0);
        if (mx < (uint32_t)kHistKeep) memcpy(R.dist.data(), ctx->h_hist_all + (size_t)i * kHistKeep, 8ull * (top + 1));
        else {
            // deeper than the bins kept per segment: count this one again with the full-width table, and add it to the totals here
            // (totals_add_kernel skipped it)
            init_accum_kernel<<<1, 1, 0, sc>>>(ctx->d_acc, nullptr);
            MSD_CUDA_OK(cudaMemsetAsync(ctx->d_hist, 0, 8ull * kDistBins, sc));
            range_stats_kernel<<<grid_for(segs[i].n_stat, 256, kSMs * 8), 256, 0, sc>>>(base + segs[i].arr_off, segs[i].n_stat, ctx->d_hist, ctx->d_acc);
            add_hist_kernel<<<kSMs, 256, 0, sc>>>(ctx->d_tot_global, ctx->d_hist, kDistBins);
            add_stats_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats, (long long)segs[i].n_stat, acc.cum_depth, acc.min_depth, acc.max_depth);
            MSD_CUDA_OK(cudaMemcpyAsync(R.dist.data(), ctx->d_hist, 8ull * (top + 1), cudaMemcpyDeviceToHost, sc));
            MSD_CUDA_OK(cudaStreamSynchronize(sc));
            ctx->launches += 4;
        }
        while (R.dist.size() > 1 && R.dist.back() == 0) R.dist.pop_back();
    }
    return 0;
}

static int check_tid(msd_ctx* ctx, int tid, const char* who) {
    if (!ctx) return MSD_EINVAL;
    if (!ctx->ran) return ctx_fail(ctx, MSD_EINVAL, "%s: call msd_run first", who);
    if (tid < 0 || tid >= ctx->n_targets || !ctx->want[tid]) return ctx_fail(ctx, MSD_EINVAL, "%s: contig %d not selected", who, tid);
    if (ctx->ranged[tid]) {
        if (ctx->pending[tid]) return ctx_fail(ctx, MSD_EINVAL, "%s: contig %d is a shard: call msd_contig_finalize first", who, tid);
        if (strcmp(who, "msd_windows") && strcmp(who, "msd_contig_stats")) return ctx_fail(ctx, MSD_EINVAL, "%s is not available on a shard of a contig (msd_set_contig_range)", who);
    } else if (ctx->n_prefilter[tid] <= 0) return ctx_fail(ctx, MSD_EINVAL, "%s: contig %d has no records (coverage() returned -2)", who, tid);
    cudaSetDevice(ctx->device);
    return 0;
}


static void fetch_region_accum(msd_ctx* ctx, msd_depth_stat* out, RegionAccum* h) {
    cudaMemcpy(h, ctx->d_racc, sizeof *h, cudaMemcpyDeviceToHost);
    if (out) { out->cum_length = (int64_t)h->cum_length; out->cum_depth = h->cum_depth; out->min_depth = h->min_depth; out->max_depth = h->max_depth; }
}


template <class F>
static int compute_runs(msd_ctx* ctx, const int32_t* arr, uint32_t n, int32_t end, F f, uint64_t* n_runs_out, int32_t keep_labels = -1) {
    cudaStream_t sc = ctx->s_compute;
    const uint32_t n_tiles = (n + kRunTile - 1) / kRunTile;
    int rc;
    if (n_tiles > ctx->run_tile_cap) {
        if ((rc = dev_alloc(ctx, &ctx->d_tile_count, (uint64_t)n_tiles))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_tile_base, (uint64_t)n_tiles))) return rc;
        ctx->run_tile_cap = n_tiles;
    }
    *n_runs_out = 0;
    if (n == 0) return 0;
    const bool trace = getenv("MSD_TRACE") != nullptr;
    const auto w0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(); };
    runs_count_kernel<F><<<n_tiles, kRunThreads, 0, sc>>>(arr, n, f, ctx->d_tile_count);
    runs_scan_kernel<<<1, 1024, 0, sc>>>(ctx->d_tile_count, n_tiles, ctx->d_tile_base, ctx->d_run_total);
    unsigned long long total = 0;
    MSD_CUDA_OK(cudaMemcpyAsync(&total, ctx->d_run_total, 8, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    const double w_count = since();
    if (total > ctx->d_run_cap) {
        const uint64_t cap = total + total / 8 + 1024;
        if ((rc = dev_alloc(ctx, &ctx->d_run_start, cap))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_run_stop, cap))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_run_val, cap))) return rc;
        ctx->d_run_cap = cap;
    }
    if (total > ctx->h_run_cap) {
        const uint64_t cap = total + total / 8 + 1024;
        if (ctx->h_run_start) cudaFreeHost(ctx->h_run_start);
        if (ctx->h_run_stop) cudaFreeHost(ctx->h_run_stop);
        if (ctx->h_run_val) cudaFreeHost(ctx->h_run_val);
        ctx->h_run_start = ctx->h_run_stop = ctx->h_run_val = nullptr; ctx->h_run_cap = 0;
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_run_start, cap * 4));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_run_stop, cap * 4));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_run_val, cap * 4));
        ctx->h_run_cap = cap;
    }
    const double w_alloc = since();
    runs_write_kernel<F><<<n_tiles, kRunThreads, 0, sc>>>(arr, n, f, ctx->d_tile_base, ctx->d_run_start, ctx->d_run_val);
    runs_stops_kernel<<<grid_for(total, 256, kSMs * 8), 256, 0, sc>>>(ctx->d_run_start, total, end, ctx->d_run_stop);
    ctx->launches += 4;
    MSD_CUDA_OK(cudaGetLastError());
    if (keep_labels >= 0 && total > 0) {
        // K9's filter (bins without a label are not yielded, :136,:140): order-preserving compaction of the runs on the device
        const uint32_t kt = (uint32_t)((total + kRunTile - 1) / kRunTile);
        if (kt > ctx->run_tile_cap) {
            if ((rc = dev_alloc(ctx, &ctx->d_tile_count, (uint64_t)kt))) return rc;
            if ((rc = dev_alloc(ctx, &ctx->d_tile_base, (uint64_t)kt))) return rc;
            ctx->run_tile_cap = kt;
        }
        if (ctx->d_keep_cap < ctx->d_run_cap) {
            if ((rc = dev_alloc(ctx, &ctx->d_keep_start, ctx->d_run_cap))) return rc;
            if ((rc = dev_alloc(ctx, &ctx->d_keep_stop, ctx->d_run_cap))) return rc;
            if ((rc = dev_alloc(ctx, &ctx->d_keep_val, ctx->d_run_cap))) return rc;
            ctx->d_keep_cap = ctx->d_run_cap;
        }
        keep_count_kernel<<<kt, kRunThreads, 0, sc>>>(ctx->d_run_val, total, keep_labels, ctx->d_tile_count);
        runs_scan_kernel<<<1, 1024, 0, sc>>>(ctx->d_tile_count, kt, ctx->d_tile_base, ctx->d_run_total);
        keep_write_kernel<<<kt, kRunThreads, 0, sc>>>(ctx->d_run_start, ctx->d_run_stop, ctx->d_run_val, total, keep_labels, ctx->d_tile_base, ctx->d_keep_start, ctx->d_keep_stop, ctx->d_keep_val);
        unsigned long long kept = 0;
        MSD_CUDA_OK(cudaMemcpyAsync(&kept, ctx->d_run_total, 8, cudaMemcpyDeviceToHost, sc));
        MSD_CUDA_OK(cudaStreamSynchronize(sc));
        std::swap(ctx->d_run_start, ctx->d_keep_start); std::swap(ctx->d_run_stop, ctx->d_keep_stop); std::swap(ctx->d_run_val, ctx->d_keep_val);   // same capacity: the kept runs are now "the runs"
        total = kept;
        ctx->launches += 3;
        MSD_CUDA_OK(cudaGetLastError());
    }
    if (trace) MSD_CUDA_OK(cudaStreamSynchronize(sc));
    const double w_write = since();
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_run_start, ctx->d_run_start, total * 4, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_run_stop, ctx->d_run_stop, total * 4, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_run_val, ctx->d_run_val, total * 4, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    if (trace) fprintf(stderr, "[msd trace] runs: %u cells -> %llu runs; count+scan %.3f ms, buffers %.3f, write+stops%s %.3f, copy to host (%.1f MB) %.3f\n", n, total, w_count, w_alloc - w_count,
                       keep_labels >= 0 ? "+keep" : "", w_write - w_alloc, total * 12 / 1e6, since() - w_write);
    *n_runs_out = total;
    return 0;
}


}  // namespace

extern "C" {

int msd_abi_version(void) { return MSD_ABI_VERSION; }

int msd_create(int device, msd_ctx** out) {
    if (!out) return MSD_EINVAL;
    *out = nullptr;
    int n = 0;
    const bool trace = getenv("MSD_TRACE") != nullptr;
    const auto c0 = std::chrono::steady_clock::now(); double c_last = 0;
    auto lap = [&](const char* what) {
        if (!trace) return;
        const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
        fprintf(stderr, "[msd trace] create: %-26s %9.3f ms\n", what, now - c_last); c_last = now;
    };
    cudaError_t e = cudaGetDeviceCount(&n);
    lap("driver init (device count)");
    if (e != cudaSuccess || n <= 0) return ctx_fail(nullptr, MSD_ENODEV, "no CUDA device: %s (libmosdepth_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return ctx_fail(nullptr, MSD_ENODEV, "device %d out of range (%d devices)", device, n);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return ctx_fail(nullptr, MSD_ENODEV, "cudaGetDeviceProperties failed");
    if (p.major != 10) return ctx_fail(nullptr, MSD_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, p.major, p.minor);
    lap("device properties");
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return ctx_fail(nullptr, MSD_ENODEV, "cudaSetDevice failed");
    lap("context");
    msd_ctx* ctx = new msd_ctx();
    ctx->device = device;
    { const uint64_t ns = env_u64("MSD_SLOTS", 4); ctx->n_slots = ns >= 8 ? 8 : ns >= 4 ? 4 : 2; }
    // MSD_PRIO=1 (default; r1n: 80.3 -> 73.3 ms per step on 176 k blocks): framing / records / scans (short kernels that release a chunk slot) are scheduled ahead of the persistent
    // inflate CTAs whenever an SM has room, instead of waiting behind them
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio_compute = env_u64("MSD_PRIO", 1) ? prio_hi : prio_lo;
    if (cudaStreamCreateWithPriority(&ctx->s_compute, cudaStreamNonBlocking, prio_compute) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx; return ctx_fail(nullptr, MSD_ENODEV, "cudaStreamCreate failed");
    }
    for (int k = 0; k < ctx->n_slots; k++) if (cudaStreamCreateWithPriority(&ctx->s_inflate[k], cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete ctx; return ctx_fail(nullptr, MSD_ENODEV, "cudaStreamCreate failed"); }
    cudaFuncSetAttribute(huffman_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecSmem);
    cudaFuncSetAttribute(huffman_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(bgzf_crc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrcSmem);
    cudaFuncSetAttribute(inflate_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
    cudaFuncSetAttribute(inflate_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    lap("streams + kernel attributes");
    if (crc_upload_tables() != cudaSuccess) { delete ctx; return ctx_fail(nullptr, MSD_ENODEV, "cannot upload the CRC tables"); }
    lap("CRC tables (module load)");
    // default filters = mosdepth defaults (-F 1796, -Q 0, no length limits, default mode)
    ctx->fp.mapq = 0; ctx->fp.min_len = -1; ctx->fp.max_len = INT64_MAX; ctx->fp.eflag = 1796; ctx->fp.iflag = 0; ctx->fp.mode = MSD_MODE_DEFAULT; ctx->fp.n_rg = 0; ctx->fp.rg_names = nullptr;
    *out = ctx;
    return MSD_OK;
}

void msd_destroy(msd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    msd_comm_destroy(ctx);
    release_input(ctx);
    dev_free(&ctx->d_rg); dev_free(&ctx->d_arena); dev_free(&ctx->d_depth_off); dev_free(&ctx->d_tlen); dev_free(&ctx->d_nprefilter); dev_free(&ctx->d_range_lo); dev_free(&ctx->d_range_hi); dev_free(&ctx->d_max_end); dev_free(&ctx->d_n_beyond); dev_free(&ctx->d_min_need);
    free_chunk_buffers(ctx);
    for (int k = 0; k < ctx->n_slots; k++) if (ctx->s_inflate[k]) cudaStreamDestroy(ctx->s_inflate[k]);
    for (int k = 0; k < 2; k++) { dev_free(&ctx->live[k].e); dev_free(&ctx->live[k].arena); dev_free(&ctx->live[k].n); dev_free(&ctx->live[k].arena_used); }
    dev_free(&ctx->d_guess); dev_free(&ctx->d_seg);
    dev_free(&ctx->d_bstart); dev_free(&ctx->d_land); dev_free(&ctx->d_count); dev_free(&ctx->d_first); dev_free(&ctx->d_cnt);
    dev_free(&ctx->d_recbase); dev_free(&ctx->d_recoff); dev_free(&ctx->d_status); dev_free(&ctx->d_ticket); dev_free(&ctx->d_inflate_err);
    dev_free(&ctx->d_cs); dev_free(&ctx->d_rc);
    dev_free(&ctx->mc.cls); dev_free(&ctx->mc.qhash); dev_free(&ctx->mc.endpos); dev_free(&ctx->mc.next); dev_free(&ctx->mc.slot);
    dev_free(&ctx->mt.word); dev_free(&ctx->mt.head); dev_free(&ctx->mt.live);
    dev_free(&ctx->d_tile_state); dev_free(&ctx->d_scan_ticket); dev_free(&ctx->d_hist); dev_free(&ctx->d_acc); dev_free(&ctx->d_racc);
    dev_free(&ctx->d_dist512); dev_free(&ctx->d_rdist); dev_free(&ctx->d_tot_global); dev_free(&ctx->d_tot_region); dev_free(&ctx->d_tot_stats);
    dev_free(&ctx->d_keep_start); dev_free(&ctx->d_keep_stop); dev_free(&ctx->d_keep_val);
    dev_free(&ctx->d_run_start); dev_free(&ctx->d_run_stop); dev_free(&ctx->d_run_val); dev_free(&ctx->d_tile_count); dev_free(&ctx->d_tile_base); dev_free(&ctx->d_run_total);
    if (ctx->h_run_start) cudaFreeHost(ctx->h_run_start);
    if (ctx->h_run_stop) cudaFreeHost(ctx->h_run_stop);
    if (ctx->h_run_val) cudaFreeHost(ctx->h_run_val);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    bz_release(ctx->bz);
    dev_free(&ctx->d_acc_all); dev_free(&ctx->d_segs); dev_free(&ctx->d_hist_all);
    if (ctx->h_acc_all) cudaFreeHost(ctx->h_acc_all); if (ctx->h_segs) cudaFreeHost(ctx->h_segs); if (ctx->h_hist_all) cudaFreeHost(ctx->h_hist_all);
    dev_free(&ctx->d_pf_values); dev_free(&ctx->d_pf_acc); dev_free(&ctx->d_pf_dist); dev_free(&ctx->d_pf_table);
    if (ctx->h_pf_values) cudaFreeHost(ctx->h_pf_values); if (ctx->h_pf_acc) cudaFreeHost(ctx->h_pf_acc); if (ctx->h_pf_dist) cudaFreeHost(ctx->h_pf_dist); if (ctx->h_pf_table) cudaFreeHost(ctx->h_pf_table);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->marks) if (e) cudaEventDestroy(e);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    delete ctx;
}

const char* msd_last_error(const msd_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int msd_open_mem(msd_ctx* ctx, const void* bam_bytes, uint64_t n_bytes, int n_targets, const uint32_t* target_len, uint64_t first_record_voff) {
    if (!ctx || !bam_bytes || n_bytes < 28) return ctx_fail(ctx, MSD_EINVAL, "msd_open_mem: bad buffer");
    cudaSetDevice(ctx->device);
    release_input(ctx);
    // chunk buffers depend on the input size: drop the old ones
    free_chunk_buffers(ctx);
    for (int k = 0; k < 2; k++) { dev_free(&ctx->live[k].e); dev_free(&ctx->live[k].arena); dev_free(&ctx->live[k].n); dev_free(&ctx->live[k].arena_used); }
    dev_free(&ctx->mc.cls); dev_free(&ctx->mc.qhash); dev_free(&ctx->mc.endpos); dev_free(&ctx->mc.next); dev_free(&ctx->mc.slot);
    dev_free(&ctx->mt.word); dev_free(&ctx->mt.head); dev_free(&ctx->mt.live); ctx->mate_ready = false;
    dev_free(&ctx->d_depth_off); dev_free(&ctx->d_tlen); dev_free(&ctx->d_nprefilter); dev_free(&ctx->d_range_lo); dev_free(&ctx->d_range_hi); dev_free(&ctx->d_max_end); dev_free(&ctx->d_n_beyond); dev_free(&ctx->d_min_need);
    ctx->bytes = (const uint8_t*)bam_bytes; ctx->n_bytes = n_bytes; ctx->mapped = false;
    return setup_common(ctx, n_targets, target_len, first_record_voff);
}

int msd_open(msd_ctx* ctx, const char* bam_path, int n_targets, const uint32_t* target_len, uint64_t first_record_voff) {
    if (!ctx || !bam_path) return ctx_fail(ctx, MSD_EINVAL, "msd_open: bad arguments");
    int fd = open(bam_path, O_RDONLY);
    if (fd < 0) return ctx_fail(ctx, MSD_EIO, "cannot open %s: %s", bam_path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 28) { close(fd); return ctx_fail(ctx, MSD_EIO, "%s: not a BAM file (too short)", bam_path); }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return ctx_fail(ctx, MSD_EIO, "mmap(%s) failed: %s", bam_path, strerror(errno));
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    int rc = msd_open_mem(ctx, m, (uint64_t)st.st_size, n_targets, target_len, first_record_voff);
    if (rc != 0) { munmap(m, (size_t)st.st_size); ctx->bytes = nullptr; return rc; }
    ctx->mapped = true;
    return MSD_OK;
}

int msd_set_spans(msd_ctx* ctx, int n, const uint64_t* voff_beg, const uint64_t* voff_end) {
    if (!ctx || n < 0 || (n > 0 && (!voff_beg || !voff_end))) return ctx_fail(ctx, MSD_EINVAL, "msd_set_spans: bad arguments");
    ctx->spans.clear();
    for (int i = 0; i < n; i++) {
        if ((voff_beg[i] >> 16) >= ctx->n_bytes) return ctx_fail(ctx, MSD_EINVAL, "span %d starts past end of file", i);
        if (voff_end[i] && voff_end[i] <= voff_beg[i]) continue;   // empty span
        ctx->spans.push_back({voff_beg[i], voff_end[i]});
    }
    if (n > 0 && ctx->spans.empty()) ctx->spans.push_back({~0ull, ~0ull});   // explicit "nothing to read"
    dev_free(&ctx->d_staged); ctx->staged_beg = ctx->staged_end = 0;
    return MSD_OK;
}

int msd_set_targets(msd_ctx* ctx, const uint8_t* want_target) {
    if (!ctx) return MSD_EINVAL;
    if (want_target) ctx->want.assign(want_target, want_target + ctx->n_targets); else ctx->want.assign(ctx->n_targets, 1);
    return MSD_OK;
}

int msd_set_filters(msd_ctx* ctx, int mapq, int64_t min_len, int64_t max_len, uint16_t eflag, uint16_t iflag,
                    const char* const* read_groups, int n_read_groups, int mode) {
    if (!ctx || mode < 0 || mode > 2 || n_read_groups < 0) return ctx_fail(ctx, MSD_EINVAL, "msd_set_filters: bad arguments");
    cudaSetDevice(ctx->device);
    ctx->fp.mapq = mapq; ctx->fp.min_len = min_len; ctx->fp.max_len = max_len; ctx->fp.eflag = eflag; ctx->fp.iflag = iflag; ctx->fp.mode = mode;
    ctx->rg_blob.clear();
    for (int i = 0; i < n_read_groups; i++) { ctx->rg_blob.append(read_groups[i]); ctx->rg_blob.push_back('\0'); }
    ctx->fp.n_rg = n_read_groups;
    dev_free(&ctx->d_rg);
    if (n_read_groups) {
        int rc = dev_alloc(ctx, &ctx->d_rg, ctx->rg_blob.size() + 1); if (rc) return rc;
        MSD_CUDA_OK(cudaMemcpy(ctx->d_rg, ctx->rg_blob.data(), ctx->rg_blob.size(), cudaMemcpyHostToDevice));
    }
    ctx->fp.rg_names = ctx->d_rg;
    return MSD_OK;
}

int msd_stage(msd_ctx* ctx) {
    if (!ctx || !ctx->bytes) return ctx_fail(ctx, MSD_EINVAL, "msd_stage: no input open");
    cudaSetDevice(ctx->device);
    uint64_t beg = ctx->first_voff >> 16, end = ctx->n_bytes;
    if (!ctx->spans.empty()) {
        beg = ~0ull; end = 0;
        for (auto& s : ctx->spans) { if (s.beg == ~0ull) continue; beg = std::min(beg, s.beg >> 16); uint64_t e = s.end ? std::min<uint64_t>((s.end >> 16) + 65536 + 64, ctx->n_bytes) : ctx->n_bytes; end = std::max(end, e); }
        if (beg == ~0ull) { beg = 0; end = 0; }
    }
    dev_free(&ctx->d_staged);
    int rc = dev_alloc(ctx, &ctx->d_staged, end - beg + 256); if (rc) return rc;
    MSD_CUDA_OK(cudaMemcpy(ctx->d_staged, ctx->bytes + beg, end - beg, cudaMemcpyHostToDevice));
    ctx->staged_beg = beg; ctx->staged_end = end;
    return MSD_OK;
}

// msd_prefetch_windows: the window loop (mosdepth.nim:720-741) of every contig that has records, in one launch; results
// go to pinned host memory and msd_windows() serves them from there.
static int run_prefetch_windows(msd_ctx* ctx) {
    ctx->pf_valid = false;
    ctx->pf_slot.assign(ctx->n_targets, -1);
    if (!ctx->pf_window) return MSD_OK;
    std::vector<WinContig> table;
    long long groups = 0, windows = 0;
    for (int t = 0; t < ctx->n_targets; t++) {
        if (!ctx->want[t] || ctx->ranged[t] || ctx->n_prefilter[t] <= 0) continue;   // contigs whose whole array msd_run has just scanned
        const long long nw = msd_n_windows(ctx->tlen[t], ctx->pf_window);
        if (nw <= 0) continue;
        WinContig c; c.arr_off = ctx->depth_off[t]; c.first_group = groups; c.first_window = windows; c.n_windows = nw; c.tlen = ctx->tlen[t]; c.pad = 0;
        ctx->pf_slot[t] = (int)table.size();
        table.push_back(c);
        groups += (nw + 31) / 32; windows += nw;
    }
    if (table.empty()) { ctx->pf_valid = true; return MSD_OK; }
    if ((uint64_t)windows * 8 > (1ull << 30)) { ctx->pf_slot.assign(ctx->n_targets, -1); return MSD_OK; }   // too many windows to keep: msd_windows computes per call
    const int nc = (int)table.size();
    if ((uint64_t)windows > ctx->pf_cap_windows) {
        dev_free(&ctx->d_pf_values); if (ctx->h_pf_values) { cudaFreeHost(ctx->h_pf_values); ctx->h_pf_values = nullptr; }
        int rc = dev_alloc(ctx, &ctx->d_pf_values, (uint64_t)windows); if (rc) return rc;
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_pf_values, (size_t)windows * 8));
        ctx->pf_cap_windows = (uint64_t)windows;
    }
    if (nc > ctx->pf_cap_contigs) {
        dev_free(&ctx->d_pf_acc); dev_free(&ctx->d_pf_dist); dev_free(&ctx->d_pf_table);
        if (ctx->h_pf_acc) cudaFreeHost(ctx->h_pf_acc); if (ctx->h_pf_dist) cudaFreeHost(ctx->h_pf_dist); if (ctx->h_pf_table) cudaFreeHost(ctx->h_pf_table);
        ctx->h_pf_acc = nullptr; ctx->h_pf_dist = nullptr; ctx->h_pf_table = nullptr;
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->d_pf_acc, (uint64_t)nc)) || (rc = dev_alloc(ctx, &ctx->d_pf_dist, (uint64_t)nc * kWindowDistBins)) || (rc = dev_alloc(ctx, &ctx->d_pf_table, (uint64_t)nc))) return rc;
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_pf_acc, sizeof(RegionAccum) * nc));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_pf_dist, 8ull * kWindowDistBins * nc));
        MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_pf_table, sizeof(WinContig) * nc));
        ctx->pf_cap_contigs = nc;
    }
    cudaStream_t sc = ctx->s_compute;
    memcpy(ctx->h_pf_table, table.data(), sizeof(WinContig) * nc);
    for (int i = 0; i < nc; i++) { RegionAccum& a = ctx->h_pf_acc[i]; a.cum_length = 0; a.cum_depth = 0; a.min_depth = 0xffffffffu; a.max_depth = 0; a.neg_seen = 0; a.pad = 0; }
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_pf_table, ctx->h_pf_table, sizeof(WinContig) * nc, cudaMemcpyHostToDevice, sc));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_pf_acc, ctx->h_pf_acc, sizeof(RegionAccum) * nc, cudaMemcpyHostToDevice, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_pf_dist, 0, 8ull * kWindowDistBins * nc, sc));
    const int grid = (int)std::min<long long>((groups + kWinWarps - 1) / kWinWarps, (long long)kSMs * 6);
    windows_all_kernel<<<grid, kWinWarps * 32, 0, sc>>>(ctx->d_arena, ctx->d_pf_table, nc, groups, ctx->pf_window, ctx->d_pf_values, ctx->d_pf_acc, ctx->d_pf_dist);
    MSD_CUDA_OK(cudaGetLastError());
    // sum_into (:761-764) for the window rows of every prefetched contig, here and once: msd_windows() then answers from the host cache
    // without a launch or a synchronisation (r1: two launches + one stream sync per contig and call)
    window_totals_kernel<<<nc, 256, 0, sc>>>(ctx->d_pf_acc, ctx->d_pf_dist, nc, ctx->d_tot_region, ctx->d_tot_stats + 4);
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_pf_values, ctx->d_pf_values, (size_t)windows * 8, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_pf_acc, ctx->d_pf_acc, sizeof(RegionAccum) * nc, cudaMemcpyDeviceToHost, sc));
    MSD_CUDA_OK(cudaMemcpyAsync(ctx->h_pf_dist, ctx->d_pf_dist, 8ull * kWindowDistBins * nc, cudaMemcpyDeviceToHost, sc));
    ctx->info.n_launches += 2;
    ctx->pf_valid = true;    // the caller synchronises the stream before msd_run returns
    return MSD_OK;
}

int msd_prefetch_windows(msd_ctx* ctx, uint32_t window) {
    if (!ctx) return MSD_EINVAL;
    ctx->pf_window = window; ctx->pf_valid = false;
    return MSD_OK;
}

int msd_run(msd_ctx* ctx) {
    if (!ctx || !ctx->bytes) return ctx_fail(ctx, MSD_EINVAL, "msd_run: no input open");
    cudaSetDevice(ctx->device);
    int rc;
    const bool trace = getenv("MSD_TRACE") != nullptr;
    const auto host_t0 = std::chrono::steady_clock::now();
    double host_last = 0;
    auto host_lap = [&](const char* what) {      // MSD_TRACE: where the HOST time of a run goes (allocations of the first run, planning, the launch loop)
        if (!trace) return;
        const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        fprintf(stderr, "[msd trace] host: %-28s %9.3f ms (at %9.3f)\n", what, now - host_last, now);
        host_last = now;
    };
    if ((rc = ensure_chunk_buffers(ctx))) return rc;
    host_lap("chunk buffers");
    if ((rc = ensure_arena(ctx))) return rc;
    host_lap("depth arena");
    if ((rc = ensure_query_buffers(ctx))) return rc;
    const bool default_mode = ctx->fp.mode == MSD_MODE_DEFAULT;
    if (default_mode && (rc = ensure_mate(ctx))) return rc;
    host_lap("query + mate buffers");
    ctx->fp.n_targets = ctx->n_targets;
    ctx->ev_used = 0;
    memset(&ctx->info, 0, sizeof ctx->info);
    ctx->ran = false; ctx->pf_valid = false;
    for (auto& r : ctx->results) { r.have = false; r.dist.clear(); }
    cudaStream_t sc = ctx->s_compute, sx = ctx->s_copy;
    StageTimes times;
    cudaEvent_t ev_begin = next_event(ctx), ev_zero_end = next_event(ctx), ev_chunks_end = next_event(ctx), ev_end = next_event(ctx);

    MSD_CUDA_OK(cudaEventRecord(ev_begin, sc));
    // a run starts the genome-wide accumulators afresh: they hold what THIS run and the queries after it add (sum_into, :761-764)
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_tot_global, 0, 8ull * kDistBins, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_tot_region, 0, 8ull * kDistBins, sc));
    init_stats_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats);
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_arena, 0, ctx->arena_cells * 4, sc));                       // init, mosdepth.nim:238-248
    MSD_CUDA_OK(cudaEventRecord(ev_zero_end, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_nprefilter, 0, sizeof(unsigned long long) * ctx->n_targets, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_max_end, 0, sizeof(uint32_t) * ctx->n_targets, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_n_beyond, 0, sizeof(unsigned long long) * ctx->n_targets, sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_min_need, 0xff, sizeof(uint32_t) * ctx->n_targets, sc));
    std::fill(ctx->pending.begin(), ctx->pending.end(), 0);
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_rc, 0, sizeof(RunCounters), sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_cs, 0, sizeof(ChunkState), sc));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_inflate_err, 0, 4, sc));
    int live_cur = 0;
    if (default_mode) for (int k = 0; k < 2; k++) { MSD_CUDA_OK(cudaMemsetAsync(ctx->live[k].n, 0, 4, sc)); MSD_CUDA_OK(cudaMemsetAsync(ctx->live[k].arena_used, 0, 4, sc)); }

    {   // the inflate streams must not run ahead of the per-run resets above
        cudaEvent_t ev_init = next_event(ctx);
        MSD_CUDA_OK(cudaEventRecord(ev_init, sc));
        for (int k = 0; k < ctx->n_slots; k++) MSD_CUDA_OK(cudaStreamWaitEvent(ctx->s_inflate[k], ev_init, 0));
        MSD_CUDA_OK(cudaStreamWaitEvent(sx, ev_init, 0));
    }
    const bool check_start = ctx->any_ranged && default_mode;
    ContigTable ct{ctx->d_depth_off, ctx->d_tlen, ctx->d_range_lo, ctx->d_range_hi, ctx->any_ranged ? ctx->d_max_end : nullptr, ctx->d_n_beyond,
                   check_start ? ctx->d_min_need : nullptr};
    if (ctx->any_ranged && ctx->fp.mode == MSD_MODE_FRAGMENT) return ctx_fail(ctx, MSD_EINVAL, "msd_set_contig_range is not supported in fragment mode (a fragment starts before its read 1)");
    const uint32_t base = ctx->carry_cap;
    // spans are taken as given (the caller merges neighbouring contigs when the bytes between them are genuine
    // file bytes; a compact per-rank buffer has cut points between its ranges, so the library must not guess)
    std::vector<Span> spans = ctx->spans;
    if (spans.empty()) spans.push_back({ctx->first_voff, 0});
    std::sort(spans.begin(), spans.end(), [](const Span& a, const Span& b) { return a.beg < b.beg; });
    // when the bytes still cross PCIe: a small first chunk (the pipeline starts after ~2 ms of copy), then chunks of
    // MSD_COPY_BLOCKS so that the serial tail behind the last byte is one modest chunk
    const uint32_t ramp0 = (uint32_t)env_u64("MSD_RAMP0_BLOCKS", 6000), copy_blocks = (uint32_t)env_u64("MSD_COPY_BLOCKS", 16000);
    const bool ramp_double = env_u64("MSD_RAMP_DOUBLE", 0) != 0;   // chunk i of a copied input holds ramp0 << i blocks until copy_blocks is reached
    const uint8_t* F = ctx->bytes; const uint64_t N = ctx->n_bytes;
    const int ingest_threads = ingest_threads_default();

    // ---- plan: every BGZF block of every span (all host threads chase BSIZE at once, ingest.hpp), cut into chunks ----------------------
    struct Plan { uint32_t b0, b1; uint64_t src, comp_bytes, infl; uint32_t nb, total, first_uoff; bool first_of_span, will_copy; };
    std::vector<BlockInfo> blocks; std::vector<Plan> plans;
    {
        std::vector<BlockInfo> sb;
        for (const Span& sp : spans) {
            if (sp.beg == ~0ull) continue;
            const uint64_t beg_c = sp.beg >> 16, end_c = sp.end ? (sp.end >> 16) : N; const uint32_t end_u = sp.end ? (uint32_t)(sp.end & 0xffff) : 0;
            // blocks that start in [beg_c, stop): the block at end_c itself belongs to the span when the span ends inside it
            const uint64_t stop = (sp.end && end_u) ? std::min<uint64_t>(end_c + 1, N) : end_c;
            uint64_t bad = 0;
            if (!chase_blocks(F, N, beg_c, stop, 2 * ingest_threads, sb, &bad)) return ctx_fail(ctx, MSD_EIO, "invalid BGZF block header at file offset %llu", (unsigned long long)bad);
            size_t i = 0; bool first = true;
            while (i < sb.size()) {
                Plan P{}; P.b0 = (uint32_t)(blocks.size() + i); P.src = sb[i].coff; P.first_of_span = first; P.first_uoff = first ? (uint32_t)(sp.beg & 0xffff) : 0u;
                P.will_copy = !(ctx->d_staged && P.src >= ctx->staged_beg && P.src < ctx->staged_end);
                const uint64_t ci = plans.size();
                const uint32_t blk_limit = !P.will_copy ? ctx->max_blocks : std::min<uint32_t>(ctx->max_blocks, ci == 0 ? ramp0 : ramp_double ? std::min<uint32_t>(copy_blocks, ramp0 << std::min<uint64_t>(ci, 8)) : copy_blocks);   // (a small first chunk for resident input was tried: 5 chunks on 4 slots, 43.3 vs 39.2 ms)
                bool truncated_last = false;
                while (i < sb.size() && P.nb < blk_limit) {
                    const BlockInfo& B = sb[i];
                    if (B.isize > 65536) return ctx_fail(ctx, MSD_EIO, "BGZF block at %llu claims %u inflated bytes", (unsigned long long)B.coff, B.isize);
                    if (P.comp_bytes + B.blen > ctx->comp_cap || P.infl + B.isize > ctx->infl_cap) { if (P.nb == 0 && P.comp_bytes == 0) return ctx_fail(ctx, MSD_ENOMEM, "chunk buffers too small for one BGZF block"); break; }
                    if (B.isize > 0) P.nb++;
                    P.comp_bytes += B.blen; i++;
                    if (sp.end && B.coff == end_c) { P.total = (uint32_t)P.infl + std::min(end_u, B.isize); truncated_last = true; P.infl += B.isize; break; }   // the span ends inside this block
                    P.infl += B.isize;
                }
                if (!truncated_last) P.total = (uint32_t)P.infl;
                P.b1 = (uint32_t)(blocks.size() + i);
                if (first && P.nb) { size_t j = P.b0 - blocks.size(); while (sb[j].isize == 0) j++; if (P.first_uoff > sb[j].isize) return ctx_fail(ctx, MSD_EINVAL, "span start offset %u beyond its block", P.first_uoff); }
                if (P.nb) { plans.push_back(P); first = false; }
                if (truncated_last) break;
            }
            blocks.insert(blocks.end(), sb.begin(), sb.end());
        }
    }
    host_lap("plan (BSIZE chase)");
    // pageable input: the bytes cross a ring of small pinned pieces on a thread of their own (IngestPump), up to n_slots - 1 chunks ahead
    bool any_copy = false; for (const Plan& P : plans) any_copy |= P.will_copy;
    const bool use_pump = any_copy && !ctx->src_pinned;
    if (use_pump && !ctx->pump) ctx->pump = new IngestPump(ctx->device, sx, ingest_threads, (size_t)env_u64("MSD_BOUNCE_MB", 32) << 20, (int)std::max<uint64_t>(2, env_u64("MSD_BOUNCE_PIECES", 8)));
    host_lap("ingest pump");
    std::vector<uint64_t> ticket(plans.size(), 0); std::vector<cudaEvent_t> h2d_t0(plans.size(), nullptr), h2d_t1(plans.size(), nullptr);
    size_t enq = 0;
    auto enqueue_upto = [&](size_t limit) {      // chunk j may be copied once chunk j - n_slots has been launched (its decode kernel frees d_comp[k])
        for (; enq < std::min(limit, plans.size()); enq++) {
            const Plan& P = plans[enq];
            if (!P.will_copy) continue;
            const int k = (int)(enq % ctx->n_slots);
            CopyRequest r; r.src = F + P.src; r.dst = ctx->d_comp[k]; r.bytes = P.comp_bytes; r.wait_before = ctx->comp_free[k];
            r.t0 = h2d_t0[enq] = next_event(ctx); r.t1 = h2d_t1[enq] = next_event(ctx); r.done = ctx->h2d_done[k];
            ticket[enq] = ctx->pump->submit(r);
        }
    };

    for (size_t chunk_idx = 0; chunk_idx < plans.size(); chunk_idx++) {
        const Plan& P = plans[chunk_idx];
        {
            // ---- descriptors of one chunk ---------------------------------------------------------------------------------------
            const int k = (int)(chunk_idx % ctx->n_slots), kd = (int)(chunk_idx % kDescSets);
            StageSet& st = ctx->stage[kd];
            if (st.used) MSD_CUDA_OK(cudaEventSynchronize(st.free_ev));
            uint32_t nb = 0; uint64_t infl = 0; const uint64_t chunk_src = P.src, comp_bytes = P.comp_bytes; const uint32_t total = P.total;
            const bool first_of_span = P.first_of_span; const uint32_t first_uoff = P.first_uoff;
            for (uint32_t bi = P.b0; bi < P.b1; bi++) {
                const BlockInfo& B = blocks[bi];
                if (B.isize > 0) {
                    st.h_desc[nb].src = (B.coff - chunk_src) + 12 + B.xlen; st.h_desc[nb].clen = B.blen - B.xlen - 20; st.h_desc[nb].isize = B.isize;
                    st.h_desc[nb].dst = (uint64_t)base + infl; st.h_rel[nb] = (uint32_t)infl; nb++;
                }
                infl += B.isize;
            }
            st.h_rel[nb] = (uint32_t)infl;
            st.used = true;
            // ---- compressed bytes -> device -------------------------------------------------------------------
            cudaStream_t si = ctx->s_inflate[k];   // inflate launches alternate between the slots' streams so that chunk k+1 fills
                                                   // the SMs while the last blocks of chunk k drain (a block takes ~10 ms)
            const uint8_t* d_comp;
            if (!P.will_copy) {
                d_comp = ctx->d_staged + (chunk_src - ctx->staged_beg);
                // descriptor src offsets assume a 4-byte aligned base: fold the misalignment into src
                const uint32_t mis = (uint32_t)((uintptr_t)d_comp & 3u);
                if (mis) { d_comp -= mis; for (uint32_t i = 0; i < nb; i++) st.h_desc[i].src += mis; }
                MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->desc_free[kd], 0));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_desc[kd], st.h_desc, sizeof(BlockDesc) * nb, cudaMemcpyHostToDevice, si));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_rel[kd], st.h_rel, sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, si));
                MSD_CUDA_OK(cudaEventRecord(st.free_ev, si));
            } else if (use_pump) {
                // the descriptors go through the inflate stream (small copies; the pump's pieces in front of them on the copy engine are
                // 32 MB = ~0.6 ms each), the bytes through the pump: wait until it has handed this chunk's last piece to the copy stream
                enqueue_upto(chunk_idx + (size_t)ctx->n_slots);
                MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->desc_free[kd], 0));
                MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->crc_done[kd], 0));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_desc[kd], st.h_desc, sizeof(BlockDesc) * nb, cudaMemcpyHostToDevice, si));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_rel[kd], st.h_rel, sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, si));
                MSD_CUDA_OK(cudaEventRecord(st.free_ev, si));
                if (!ctx->pump->wait_issued(ticket[chunk_idx])) return ctx_fail(ctx, MSD_ECUDA, "%s", ctx->pump->error());
                MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->h2d_done[k], 0));
                times.h2d.push_back({h2d_t0[chunk_idx], h2d_t1[chunk_idx]});
                d_comp = ctx->d_comp[k];
            } else {
                // pinned source: everything this chunk needs goes through the copy stream in one batch, small descriptor copies first:
                // a copy engine is FIFO, so a descriptor copy queued behind the NEXT chunk's bulk copy would hold this
                // chunk's kernels back until that bulk copy ends
                MSD_CUDA_OK(cudaStreamWaitEvent(sx, ctx->desc_free[kd], 0));   // sixteen sets: chunk k-16 is long gone
                MSD_CUDA_OK(cudaStreamWaitEvent(sx, ctx->crc_done[kd], 0));    // ... and so is its CRC kernel, which reads the descriptors off the critical path
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_desc[kd], st.h_desc, sizeof(BlockDesc) * nb, cudaMemcpyHostToDevice, sx));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_rel[kd], st.h_rel, sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, sx));
                MSD_CUDA_OK(cudaEventRecord(st.free_ev, sx));       // h_desc and h_rel of this set are free again
                MSD_CUDA_OK(cudaStreamWaitEvent(sx, ctx->comp_free[k], 0));     // the decode kernel of chunk k-n_slots has read d_comp[k]
                cudaEvent_t h0 = next_event(ctx), h1 = next_event(ctx);
                MSD_CUDA_OK(cudaEventRecord(h0, sx));
                MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_comp[k], F + chunk_src, comp_bytes, cudaMemcpyHostToDevice, sx));
                MSD_CUDA_OK(cudaEventRecord(h1, sx));
                MSD_CUDA_OK(cudaEventRecord(ctx->h2d_done[k], sx));
                MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->h2d_done[k], 0));
                times.h2d.push_back({h0, h1});
                d_comp = ctx->d_comp[k];
            }
            // ---- K1 inflate ---------------------------------------------------------------------------------
            uint8_t* infl_buf = ctx->d_infl[k];
            cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx), e2 = next_event(ctx), e3 = next_event(ctx);
            cudaEvent_t f0 = next_event(ctx);
            cudaEvent_t tr0 = trace ? next_event(ctx) : nullptr;
            MSD_CUDA_OK(cudaStreamWaitEvent(si, ctx->infl_free[k], 0));      // consumers of the previous contents of d_infl[k] are done
            MSD_CUDA_OK(cudaMemsetAsync(ctx->d_ticket + 4 * k, 0, 16, si));
            MSD_CUDA_OK(cudaEventRecord(e0, si));
            launch_inflate(si, d_comp, infl_buf, ctx->d_desc[kd], (int)nb, ctx->d_ticket + 4 * k, ctx->d_tokens[k], ctx->d_ntok[k],
                           ctx->d_status + (size_t)k * ctx->max_blocks, ctx->d_inflate_err, ctx->d_crc[k], ctx->check_crc, ctx->comp_free[k], ctx->infl_done[k], tr0, ctx->d_ring[k]);
            ctx->info.n_launches += (ctx->check_crc ? 1 : 0) + (inflate_fused_enabled() ? 1 : 2); ctx->info.n_inflate_launches++;
            MSD_CUDA_OK(cudaEventRecord(e1, si));
            MSD_CUDA_OK(cudaEventRecord(ctx->crc_done[kd], si));
            MSD_CUDA_OK(cudaStreamWaitEvent(sc, ctx->infl_done[k], 0));
            MSD_CUDA_OK(cudaEventRecord(f0, sc));
            // ---- K2 framing ---------------------------------------------------------------------------------
            const int nblk = (int)nb + 1;   // + carry pseudo-block
            frame_setup_kernel<<<(nb + 255) / 256, 256, 0, sc>>>(ctx->d_cs, ctx->d_rel[kd], (int)nb, base, first_of_span ? first_uoff : 0u, total, first_of_span ? 1 : 0, ctx->d_bstart);
            frame_guess_kernel<<<(nblk * 32 + 255) / 256, 256, 0, sc>>>(infl_buf, ctx->d_cs, ctx->d_bstart, nblk, ctx->n_targets, ctx->d_guess);
            frame_segments_kernel<<<1, 1024, 0, sc>>>(ctx->d_cs, ctx->d_guess, nblk, ctx->d_seg);
            frame_spec_kernel<<<(nblk + 127) / 128, 128, 0, sc>>>(infl_buf, ctx->d_cs, ctx->d_seg, nblk, ctx->d_land, ctx->d_count);
            frame_link_kernel<<<1, 1024, 0, sc>>>(infl_buf, ctx->d_cs, ctx->d_seg, nblk, ctx->d_land, ctx->d_count, ctx->d_first, ctx->d_cnt, ctx->d_recbase);
            frame_emit_kernel<<<(nblk + 127) / 128, 128, 0, sc>>>(infl_buf, ctx->d_cs, ctx->d_first, ctx->d_cnt, ctx->d_recbase, nblk, ctx->d_recoff);
            ctx->info.n_launches += 6;
            MSD_CUDA_OK(cudaEventRecord(e2, sc));
            // ---- K3/K5 (+K4) ------------------------------------------------------------------------------------
            const int rgrid = kSMs * 8;
            records_kernel<<<rgrid, 256, 0, sc>>>(infl_buf, ctx->d_recoff, ctx->d_cs, ctx->fp, ct, ctx->d_arena, ctx->d_nprefilter, ctx->mc, ctx->d_rc);
            ctx->info.n_launches++;
            if (default_mode) {
                LiveList& L = ctx->live[live_cur]; LiveList& Ln = ctx->live[live_cur ^ 1];
                MSD_CUDA_OK(cudaMemsetAsync(ctx->mt.word, 0xff, 8ull * ctx->mate_slots, sc));
                MSD_CUDA_OK(cudaMemsetAsync(ctx->mt.head, 0xff, 4ull * ctx->mate_slots, sc));
                MSD_CUDA_OK(cudaMemsetAsync(ctx->mt.live, 0xff, 4ull * ctx->mate_slots, sc));
                MSD_CUDA_OK(cudaMemsetAsync(Ln.n, 0, 4, sc));
                MSD_CUDA_OK(cudaMemsetAsync(Ln.arena_used, 0, 4, sc));
                mate_reinsert_kernel<<<kSMs * 2, 256, 0, sc>>>(ctx->mt, L, infl_buf, ctx->d_recoff, ctx->d_rc);
                mate_link_kernel<<<rgrid, 256, 0, sc>>>(ctx->mt, L, infl_buf, ctx->d_recoff, ctx->d_cs, ctx->mc, 0, ctx->d_rc, ct);
                mate_link_kernel<<<rgrid, 256, 0, sc>>>(ctx->mt, L, infl_buf, ctx->d_recoff, ctx->d_cs, ctx->mc, 1, ctx->d_rc, ct);
                mate_replay_kernel<<<rgrid, 256, 0, sc>>>(ctx->mt, L, Ln, infl_buf, ctx->d_recoff, ctx->d_cs, ctx->mc, ct, ctx->d_arena, ctx->d_rc);
                mate_carry_kernel<<<kSMs * 2, 256, 0, sc>>>(L, Ln, ctx->d_rc);
                ctx->info.n_launches += 5;
                live_cur ^= 1;
            }
            MSD_CUDA_OK(cudaEventRecord(e3, sc));
            // unfinished tail -> front of the other inflated buffer
            carry_copy_kernel<<<64, 256, 0, sc>>>(infl_buf, ctx->d_infl[(k + 1) % ctx->n_slots], ctx->d_cs, base, ctx->carry_cap, &ctx->d_rc->error);
            ctx->info.n_launches++;
            MSD_CUDA_OK(cudaEventRecord(ctx->desc_free[kd], sc));  // d_desc[kd], d_rel[kd] may be overwritten from here on (d_comp[k]: after the decode kernel)
            MSD_CUDA_OK(cudaEventRecord(ctx->infl_free[k], sc));   // ... and so may d_infl[k]
            MSD_CUDA_OK(cudaGetLastError());
            times.inflate.push_back({e0, e1}); times.frame.push_back({f0, e2}); times.records.push_back({e2, e3}); times.k1.push_back({tr0, f0});
            ctx->info.bytes_compressed += comp_bytes; ctx->info.bytes_inflated += infl; ctx->info.n_blocks += nb;
        }
    }
    MSD_CUDA_OK(cudaEventRecord(ev_chunks_end, sc));
    host_lap("chunk loop (launches)");
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    for (int k = 0; k < ctx->n_slots; k++) MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_inflate[k]));   // the CRC kernels run beside framing and records
    // ---- errors raised by the kernels --------------------------------------------------------------------------
    uint32_t inflate_err = 0; ChunkState cs_h; RunCounters rc_h;
    MSD_CUDA_OK(cudaMemcpy(&inflate_err, ctx->d_inflate_err, 4, cudaMemcpyDeviceToHost));
    MSD_CUDA_OK(cudaMemcpy(&cs_h, ctx->d_cs, sizeof cs_h, cudaMemcpyDeviceToHost));
    MSD_CUDA_OK(cudaMemcpy(&rc_h, ctx->d_rc, sizeof rc_h, cudaMemcpyDeviceToHost));
    if (inflate_err) return ctx_fail(ctx, MSD_EDATA, "%u BGZF block(s) failed to inflate (corrupt DEFLATE stream)", inflate_err);
    if (cs_h.error) return ctx_fail(ctx, MSD_EDATA, "corrupt BAM record chain (bad block_size)");
    if (rc_h.error == 1) return ctx_fail(ctx, MSD_ENOMEM, "mate hash table full: raise MSD_MATE_SLOTS");
    if (rc_h.error == 2) return ctx_fail(ctx, MSD_ENOMEM, "live mate list full: raise MSD_LIVE_CAP or unfinished record larger than MSD_CARRY_MB");
    if (rc_h.error == 3) return ctx_fail(ctx, MSD_ENOMEM, "live mate arena full: raise MSD_LIVE_ARENA_MB");
    ctx->info.n_events = rc_h.n_events; ctx->info.n_records = rc_h.n_records;
    // ---- per contig: prefix sum + whole-contig stats -----------------------------------------------------------
    std::vector<unsigned long long> npre(ctx->n_targets);
    MSD_CUDA_OK(cudaMemcpy(npre.data(), ctx->d_nprefilter, sizeof(unsigned long long) * ctx->n_targets, cudaMemcpyDeviceToHost));
    host_lap("wait for the device");
    cudaEvent_t s0 = next_event(ctx), s1 = next_event(ctx);
    MSD_CUDA_OK(cudaEventRecord(s0, sc));
    std::vector<unsigned long long> nbey(ctx->n_targets, 0);
    if (ctx->any_ranged) {
        MSD_CUDA_OK(cudaMemcpy(ctx->max_end.data(), ctx->d_max_end, sizeof(uint32_t) * ctx->n_targets, cudaMemcpyDeviceToHost));
        MSD_CUDA_OK(cudaMemcpy(nbey.data(), ctx->d_n_beyond, sizeof(unsigned long long) * ctx->n_targets, cudaMemcpyDeviceToHost));
        ctx->min_need.assign(ctx->n_targets, 0xffffffffu);
        if (check_start) {
            MSD_CUDA_OK(cudaMemcpy(ctx->min_need.data(), ctx->d_min_need, sizeof(uint32_t) * ctx->n_targets, cudaMemcpyDeviceToHost));
            for (int t = 0; t < ctx->n_targets; t++) {
                const uint32_t cover = std::min(ctx->covered_from[t], ctx->range_lo[t]);   // never declared: only the owned range is known to be there
                if (ctx->want[t] && ctx->ranged[t] && ctx->range_lo[t] && ctx->min_need[t] < cover)
                    return ctx_fail(ctx, MSD_ESPAN, "contig %d: an owned read looked for a mate at position %u, but the span is only known to hold the records from %u on: extend the span start of this shard (msd_set_contig_cover)", t, ctx->min_need[t], cover);
            }
        }
    }
    // one segment per contig that had records (tlen + 1 cells scanned, tlen of them feed the stats, :746-747) or per owned range of
    // a sharded contig (prefix sum of this shard's own events over [lo, last event]; the depth it leaves beyond hi goes to the next
    // shard -- msd_exchange_spills, or msd_depth_copy / msd_depth_add -- and its stats come with msd_contig_finalize)
    std::vector<ScanSeg> segs; std::vector<int> seg_tid;
    auto flush = [&](bool last) -> int {
        if (segs.empty()) return 0;
        int r = run_segments(ctx, ctx->d_arena, segs, true); if (r) return r;
        if (last) return 0;
        MSD_CUDA_OK(cudaStreamSynchronize(sc));
        r = collect_segments(ctx, ctx->d_arena, segs, seg_tid);
        segs.clear(); seg_tid.clear();
        return r;
    };
    for (int t = 0; t < ctx->n_targets; t++) {
        ctx->n_prefilter[t] = (long long)npre[t];
        if (!ctx->want[t]) continue;
        ScanSeg sg{};
        if (ctx->ranged[t]) {
            const uint32_t lo = ctx->range_lo[t], hi = std::min(ctx->range_hi[t], ctx->tlen[t]);
            if (hi < ctx->tlen[t] && nbey[t] == 0)
                return ctx_fail(ctx, MSD_EINVAL, "contig %d: the span ends before any record at or beyond %u was seen; extend the span of this shard", t, hi);
            const uint32_t last = std::max(ctx->max_end[t], lo);
            sg.arr_off = ctx->depth_off[t] + lo; sg.n_scan = std::min<uint32_t>(last, ctx->tlen[t]) + 1 - lo; sg.n_stat = 0;
            ctx->pending[t] = 1;
            ctx->info.n_depth_cells += (uint64_t)hi - lo;
        } else {
            if (npre[t] == 0) continue;
            sg.arr_off = ctx->depth_off[t]; sg.n_scan = ctx->tlen[t] + 1; sg.n_stat = ctx->tlen[t];
            ctx->info.n_depth_cells += (uint64_t)ctx->tlen[t] + 1;
            if (ctx->tlen[t] == 0) { ContigResult& R = ctx->results[t]; R.have = true; R.stat.cum_length = 0; R.stat.cum_depth = 0; R.stat.min_depth = 0xffffffffu; R.stat.max_depth = 0; R.dist.assign(1, 0); }
        }
        segs.push_back(sg); seg_tid.push_back(t);
        if ((int)segs.size() == kMaxSegsPerLaunch && (rc = flush(false))) return rc;
    }
    if ((rc = flush(true))) return rc;
    MSD_CUDA_OK(cudaEventRecord(s1, sc));
    if ((rc = run_prefetch_windows(ctx))) return rc;
    MSD_CUDA_OK(cudaEventRecord(ev_end, sc));
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    if ((rc = collect_segments(ctx, ctx->d_arena, segs, seg_tid))) return rc;
    host_lap("scans + windows + collect");
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end); ctx->info.ms_total = ms;
    cudaEventElapsedTime(&ms, ev_begin, ev_zero_end); ctx->info.ms_zero = ms;
    cudaEventElapsedTime(&ms, s0, s1); ctx->info.ms_scan = ms;
    if (trace) {
        auto rel = [&](cudaEvent_t e) { float m = 0; cudaEventElapsedTime(&m, ev_begin, e); return m; };
        for (size_t i = 0; i < times.inflate.size(); i++)
            fprintf(stderr, "[msd trace] chunk %zu: h2d %.2f..%.2f  inflate %.2f (decoded %.2f, resolved <= %.2f) ..%.2f  frame ..%.2f  records ..%.2f\n", i,
                    i < times.h2d.size() ? rel(times.h2d[i].first) : -1.f, i < times.h2d.size() ? rel(times.h2d[i].second) : -1.f,
                    rel(times.inflate[i].first), times.k1[i].first ? rel(times.k1[i].first) : -1.f, rel(times.k1[i].second), rel(times.inflate[i].second), rel(times.frame[i].second), rel(times.records[i].second));
        fprintf(stderr, "[msd trace] chunks end %.2f  scans %.2f..%.2f  end %.2f\n", rel(ev_chunks_end), rel(s0), rel(s1), rel(ev_end));
    }
    if (!times.inflate.empty()) { float sp = 0; cudaEventElapsedTime(&sp, times.inflate.front().first, times.inflate.back().second); ctx->info.ms_inflate_span = sp; }
    ctx->info.ms_h2d = sum_ms(times.h2d); ctx->info.ms_inflate = sum_ms(times.inflate); ctx->info.ms_frame = sum_ms(times.frame); ctx->info.ms_records = sum_ms(times.records);
    ctx->ran = true;
    ctx->launches += ctx->info.n_launches;
    return MSD_OK;
}

int msd_last_run_info(const msd_ctx* ctx, msd_run_info* out) { if (!ctx || !out) return MSD_EINVAL; *out = ctx->info; return MSD_OK; }

// ---- intra-contig sharding ------------------------------------------------------------------------------------------------
int msd_set_contig_range(msd_ctx* ctx, int tid, uint32_t lo, uint32_t hi) {
    if (!ctx || tid < 0 || tid >= ctx->n_targets) return ctx_fail(ctx, MSD_EINVAL, "msd_set_contig_range: bad contig");
    const uint32_t tlen = ctx->tlen[tid];
    if (hi == 0 || hi > tlen) hi = tlen;
    if (lo == 0 && hi == tlen) { ctx->ranged[tid] = 0; ctx->range_lo[tid] = 0; ctx->range_hi[tid] = 0xffffffffu; }
    else {
        if (lo >= hi || (lo & 31u)) return ctx_fail(ctx, MSD_EINVAL, "msd_set_contig_range: need lo < hi and lo a multiple of 32");
        ctx->ranged[tid] = 1; ctx->range_lo[tid] = lo; ctx->range_hi[tid] = hi == tlen ? 0xffffffffu : hi;
    }
    ctx->any_ranged = false;
    for (int t = 0; t < ctx->n_targets; t++) ctx->any_ranged |= ctx->ranged[t] != 0;
    ctx->ran = false;
    return MSD_OK;
}

int msd_set_contig_cover(msd_ctx* ctx, int tid, uint32_t covered_from) {
    if (!ctx || tid < 0 || tid >= ctx->n_targets) return ctx_fail(ctx, MSD_EINVAL, "msd_set_contig_cover: bad contig");
    ctx->covered_from[tid] = covered_from;
    return MSD_OK;
}

int msd_contig_span_need(msd_ctx* ctx, int tid, uint32_t* need_from, uint32_t* covered_from) {
    if (!ctx || tid < 0 || tid >= ctx->n_targets) return ctx_fail(ctx, MSD_EINVAL, "msd_contig_span_need: bad contig");
    if (need_from) *need_from = ctx->min_need[tid];
    if (covered_from) *covered_from = std::min(ctx->covered_from[tid], ctx->range_lo[tid]);
    return MSD_OK;
}

static int check_pending(msd_ctx* ctx, int tid, const char* who) {
    if (!ctx || tid < 0 || tid >= ctx->n_targets || !ctx->ranged[tid] || !ctx->pending[tid]) return ctx_fail(ctx, MSD_EINVAL, "%s: contig is not a shard waiting for msd_contig_finalize()", who);
    cudaSetDevice(ctx->device);
    return MSD_OK;
}

int msd_contig_spill(msd_ctx* ctx, int tid, uint32_t* from, uint32_t* n) {
    int rc = check_pending(ctx, tid, "msd_contig_spill"); if (rc) return rc;
    const uint32_t tlen = ctx->tlen[tid], hi = std::min(ctx->range_hi[tid], tlen);
    if (from) *from = hi;
    if (n) *n = ctx->max_end[tid] > hi ? std::min(ctx->max_end[tid], tlen) - hi : 0;
    return MSD_OK;
}

int msd_depth_copy(msd_ctx* ctx, int tid, uint32_t pos, uint32_t n, int32_t* dst, int dst_is_device) {
    int rc = check_pending(ctx, tid, "msd_depth_copy"); if (rc) return rc;
    if (!dst || (uint64_t)pos + n > (uint64_t)ctx->tlen[tid] + 1) return ctx_fail(ctx, MSD_EINVAL, "msd_depth_copy: bad range");
    if (n) MSD_CUDA_OK(cudaMemcpyAsync(dst, ctx->d_arena + ctx->depth_off[tid] + pos, 4ull * n, dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->s_compute));
    MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
    return MSD_OK;
}

int msd_depth_add(msd_ctx* ctx, int tid, uint32_t pos, uint32_t n, const int32_t* src, int src_is_device) {
    int rc = check_pending(ctx, tid, "msd_depth_add"); if (rc) return rc;
    if (!src || (uint64_t)pos + n > (uint64_t)ctx->tlen[tid] + 1) return ctx_fail(ctx, MSD_EINVAL, "msd_depth_add: bad range");
    if (!n) return MSD_OK;
    const int32_t* d_src = src;
    if (!src_is_device) {
        if ((rc = ensure_scratch(ctx, 4ull * n))) return rc;
        MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_scratch, src, 4ull * n, cudaMemcpyHostToDevice, ctx->s_compute));
        d_src = (const int32_t*)ctx->d_scratch;
    }
    add_cells_kernel<<<grid_for(n, 256, kSMs * 8), 256, 0, ctx->s_compute>>>(ctx->d_arena + ctx->depth_off[tid] + pos, d_src, n);
    ctx->launches++;
    MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
    return MSD_OK;
}

// stats + dist of the owned ranges of the given shards (all of them in one launch), after their neighbours' spills were added
static int finalize_shards(msd_ctx* ctx, const std::vector<int>& tids) {
    int rc;
    if ((rc = ensure_query_buffers(ctx))) return rc;
    for (size_t b0 = 0; b0 < tids.size(); b0 += kMaxSegsPerLaunch) {
        std::vector<ScanSeg> segs; std::vector<int> seg_tid;
        for (size_t i = b0; i < std::min(tids.size(), b0 + (size_t)kMaxSegsPerLaunch); i++) {
            const int t = tids[i];
            const uint32_t tlen = ctx->tlen[t], lo = ctx->range_lo[t], hi = std::min(ctx->range_hi[t], tlen);
            ScanSeg sg{}; sg.arr_off = ctx->depth_off[t] + lo; sg.n_scan = hi - lo; sg.n_stat = hi - lo;
            segs.push_back(sg); seg_tid.push_back(t);
        }
        if ((rc = run_segments(ctx, ctx->d_arena, segs, false))) return rc;
        MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
        if ((rc = collect_segments(ctx, ctx->d_arena, segs, seg_tid))) return rc;
        ctx->launches += 3;
    }
    for (int t : tids) ctx->pending[t] = 0;
    return MSD_OK;
}

int msd_contig_finalize(msd_ctx* ctx, int tid) {
    int rc = check_pending(ctx, tid, "msd_contig_finalize"); if (rc) return rc;
    return finalize_shards(ctx, std::vector<int>(1, tid));
}

int msd_contig_records(msd_ctx* ctx, int tid, int64_t* n) {
    if (!ctx || !ctx->ran) return ctx_fail(ctx, MSD_EINVAL, "msd_contig_records: call msd_run first");
    if (tid < 0 || tid >= ctx->n_targets || !ctx->want[tid]) { if (n) *n = 0; return -1; }
    if (n) *n = ctx->n_prefilter[tid];
    return ctx->n_prefilter[tid] > 0 ? tid : -2;
}

int msd_contig_stats(msd_ctx* ctx, int tid, msd_depth_stat* chrom_stat, int64_t* dist, int64_t dist_cap, int64_t* dist_len) {
    int rc = check_tid(ctx, tid, "msd_contig_stats"); if (rc) return rc;
    const ContigResult& R = ctx->results[tid];
    if (chrom_stat) *chrom_stat = R.stat;
    if (dist_len) *dist_len = (int64_t)R.dist.size();
    if (dist) {
        if (dist_cap < (int64_t)R.dist.size()) return ctx_fail(ctx, MSD_EINVAL, "msd_contig_stats: dist_cap %lld < %zu", (long long)dist_cap, R.dist.size());
        memcpy(dist, R.dist.data(), 8 * R.dist.size());
    }
    return MSD_OK;
}

int64_t msd_n_windows(uint32_t tlen, uint32_t window) {
    // window_gen (mosdepth.nim:382-388): [k*w, (k+1)*w) while start + w < tlen, then [start, tlen) if start != tlen
    if (tlen == 0) return 0;
    if (window == 0) return -1;
    return ((int64_t)tlen + window - 1) / window;
}

int64_t msd_windows(msd_ctx* ctx, int tid, uint32_t window, int use_median, double* value_out, int64_t value_cap,
                    msd_depth_stat* region_stat, int64_t* region_dist512) {
    int rc = check_tid(ctx, tid, "msd_windows"); if (rc) return rc;
    if (window == 0) return ctx_fail(ctx, MSD_EINVAL, "msd_windows: window must be > 0");
    // a shard [lo, hi) of a contig is the array arr + lo of length hi - lo: lo is a multiple of the window and hi is one too
    // or the contig end, so window_gen (:382-388) over the shard yields exactly the contig's windows that start in it
    const uint32_t lo = ctx->ranged[tid] ? ctx->range_lo[tid] : 0;
    const uint32_t tlen = ctx->ranged[tid] ? std::min(ctx->range_hi[tid], ctx->tlen[tid]) - lo : ctx->tlen[tid];
    if (ctx->ranged[tid] && (lo % window || (tlen % window && lo + tlen != ctx->tlen[tid]))) return ctx_fail(ctx, MSD_EINVAL, "msd_windows: the shard [%u, %u) does not start and end on multiples of the window %u", lo, lo + tlen, window);
    const int64_t nw = msd_n_windows(tlen, window);
    if (value_out && value_cap < nw) return ctx_fail(ctx, MSD_EINVAL, "msd_windows: value_cap %lld < %lld windows", (long long)value_cap, (long long)nw);
    if (ctx->pf_valid && window == ctx->pf_window && !use_median && nw > 0 && tid < (int)ctx->pf_slot.size() && ctx->pf_slot[tid] >= 0) {
        // computed by msd_run (msd_prefetch_windows), totals included: serve from the host cache
        const int i = ctx->pf_slot[tid];
        const RegionAccum& h = ctx->h_pf_acc[i];
        if (region_stat) { region_stat->cum_length = (int64_t)h.cum_length; region_stat->cum_depth = h.cum_depth; region_stat->min_depth = h.min_depth; region_stat->max_depth = h.max_depth; }
        if (value_out) memcpy(value_out, ctx->h_pf_values + ctx->h_pf_table[i].first_window, (size_t)nw * 8);
        if (region_dist512) memcpy(region_dist512, ctx->h_pf_dist + (size_t)i * kWindowDistBins, 8 * kWindowDistBins);
        return nw;
    }
    if ((rc = ensure_scratch(ctx, (uint64_t)std::max<int64_t>(nw, 1) * 8))) return rc;
    cudaStream_t sc = ctx->s_compute;
    init_accum_kernel<<<1, 1, 0, sc>>>(nullptr, ctx->d_racc);
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_dist512, 0, 8 * kWindowDistBins, sc));
    if (nw > 0) {
        if (use_median)
            windows_kernel<<<grid_for((uint64_t)nw, 256, kSMs * 8), 256, 0, sc>>>(ctx->d_arena + ctx->depth_off[tid] + lo, tlen, window, nw, use_median,
                                                                                (double*)ctx->d_scratch, ctx->d_racc, ctx->d_dist512);
        else
            windows_mean_kernel<<<grid_for((uint64_t)nw, kWinWarps * 32, kSMs * 4), kWinWarps * 32, 0, sc>>>(ctx->d_arena + ctx->depth_off[tid] + lo, tlen, window, nw,
                                                                                                     (double*)ctx->d_scratch, ctx->d_racc, ctx->d_dist512);
        MSD_CUDA_OK(cudaGetLastError());
    }
    add_hist_kernel<<<2, 256, 0, sc>>>(ctx->d_tot_region, ctx->d_dist512, kWindowDistBins);
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    ctx->launches += 4;
    RegionAccum h; fetch_region_accum(ctx, region_stat, &h);
    if (h.neg_seen) return ctx_fail(ctx, MSD_ENEGDEPTH, "error setting negative depth value (median mode)");
    add_stats_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats + 4, (long long)h.cum_length, h.cum_depth, h.min_depth, h.max_depth);
    if (value_out && nw > 0) MSD_CUDA_OK(cudaMemcpy(value_out, ctx->d_scratch, (size_t)nw * 8, cudaMemcpyDeviceToHost));
    if (region_dist512) MSD_CUDA_OK(cudaMemcpy(region_dist512, ctx->d_dist512, 8 * kWindowDistBins, cudaMemcpyDeviceToHost));
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    return nw;
}

int msd_regions(msd_ctx* ctx, int tid, int64_t n, const uint32_t* start, const uint32_t* stop, int use_median, double* value_out,
                msd_depth_stat* region_stat, int64_t* region_dist, int64_t dist_cap, int64_t* dist_len, const int32_t* thresholds,
                int n_thr, int64_t* thr_counts) {
    int rc = check_tid(ctx, tid, "msd_regions"); if (rc) return rc;
    if (n < 0 || (n > 0 && (!start || !stop)) || n_thr < 0 || (n_thr > 0 && (!thresholds || !thr_counts))) return ctx_fail(ctx, MSD_EINVAL, "msd_regions: bad arguments");
    const uint32_t tlen = ctx->tlen[tid];
    cudaStream_t sc = ctx->s_compute;
    // scratch layout: starts | stops | values | thresholds | thr_counts
    const uint64_t nn = (uint64_t)std::max<int64_t>(n, 1);
    const uint64_t o_start = 0, o_stop = o_start + nn * 4, o_val = ((o_stop + nn * 4 + 7) / 8) * 8, o_thr = o_val + nn * 8, o_cnt = ((o_thr + (uint64_t)std::max(n_thr, 1) * 4 + 7) / 8) * 8;
    const uint64_t bytes = o_cnt + nn * (uint64_t)std::max(n_thr, 1) * 8;
    if ((rc = ensure_scratch(ctx, bytes))) return rc;
    uint8_t* S = (uint8_t*)ctx->d_scratch;
    init_accum_kernel<<<1, 1, 0, sc>>>(nullptr, ctx->d_racc);
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_rdist, 0, 8ull * kDistBins, sc));
    if (n > 0) {
        MSD_CUDA_OK(cudaMemcpyAsync(S + o_start, start, (size_t)n * 4, cudaMemcpyHostToDevice, sc));
        MSD_CUDA_OK(cudaMemcpyAsync(S + o_stop, stop, (size_t)n * 4, cudaMemcpyHostToDevice, sc));
        if (n_thr) MSD_CUDA_OK(cudaMemcpyAsync(S + o_thr, thresholds, (size_t)n_thr * 4, cudaMemcpyHostToDevice, sc));
        regions_kernel<<<grid_for((uint64_t)n, kRegWarps, kSMs * 8), kRegWarps * 32, 0, sc>>>(ctx->d_arena + ctx->depth_off[tid], tlen, n, (const uint32_t*)(S + o_start),
                                                                           (const uint32_t*)(S + o_stop), use_median, (double*)(S + o_val), ctx->d_racc, ctx->d_rdist,
                                                                           (const int32_t*)(S + o_thr), n_thr, (unsigned long long*)(S + o_cnt));
        MSD_CUDA_OK(cudaGetLastError());
    }
    // sum_into for the `total_region` rows (:761-764): not when the call only asks for threshold counts over windows that
    // msd_windows() has already accounted for (region_stat == NULL)
    const bool to_totals = region_stat != nullptr;
    if (to_totals) add_hist_kernel<<<kSMs, 256, 0, sc>>>(ctx->d_tot_region, ctx->d_rdist, kDistBins);
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    ctx->launches += 4;
    RegionAccum h; fetch_region_accum(ctx, region_stat, &h);
    if (h.neg_seen) return ctx_fail(ctx, MSD_ENEGDEPTH, "error setting negative depth value (median mode)");
    if (to_totals) add_stats_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats + 4, (long long)h.cum_length, h.cum_depth, h.min_depth, h.max_depth);
    if (n > 0 && value_out) MSD_CUDA_OK(cudaMemcpy(value_out, S + o_val, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (n > 0 && n_thr) MSD_CUDA_OK(cudaMemcpy(thr_counts, S + o_cnt, (size_t)n * n_thr * 8, cudaMemcpyDeviceToHost));
    if (region_dist || dist_len) {
        // length = 1 + highest non-zero bin, bounded by the region maximum
        int64_t top = (h.max_depth > 0x7fffffffu) ? (int64_t)kMaxCoverage : (int64_t)std::min<uint32_t>(h.max_depth, kMaxCoverage);
        std::vector<int64_t> tmp((size_t)top + 1, 0);
        MSD_CUDA_OK(cudaMemcpy(tmp.data(), ctx->d_rdist, 8ull * (top + 1), cudaMemcpyDeviceToHost));
        while (tmp.size() > 1 && tmp.back() == 0) tmp.pop_back();
        if (dist_len) *dist_len = (int64_t)tmp.size();
        if (region_dist) { if (dist_cap < (int64_t)tmp.size()) return ctx_fail(ctx, MSD_EINVAL, "msd_regions: dist_cap too small"); memcpy(region_dist, tmp.data(), 8 * tmp.size()); }
    }
    MSD_CUDA_OK(cudaStreamSynchronize(sc));
    return MSD_OK;
}

int msd_per_base_runs(msd_ctx* ctx, int tid, const int32_t** start, const int32_t** stop, const int32_t** depth, int64_t* n_runs) {
    int rc = check_tid(ctx, tid, "msd_per_base_runs"); if (rc) return rc;
    if ((rc = ensure_query_buffers(ctx))) return rc;
    const uint32_t tlen = ctx->tlen[tid];
    uint64_t total = 0;
    if ((rc = compute_runs(ctx, ctx->d_arena + ctx->depth_off[tid], tlen, (int32_t)tlen, IdentityF(), &total))) return rc;
    if (start) *start = ctx->h_run_start; if (stop) *stop = ctx->h_run_stop; if (depth) *depth = ctx->h_run_val;
    if (n_runs) *n_runs = (int64_t)total;
    return MSD_OK;
}

int msd_per_base_bgzf(msd_ctx* ctx, int tid, const char* chrom, const uint8_t** bgzf, int64_t* n_bytes, const msd_bgzf_member** members, int64_t* n_members,
                      const int32_t** start, const int32_t** stop, const int32_t** depth, int64_t* n_runs) {
    int rc = check_tid(ctx, tid, "msd_per_base_bgzf"); if (rc) return rc;
    if (!chrom || !bgzf || !n_bytes || !members || !n_members) return ctx_fail(ctx, MSD_EINVAL, "msd_per_base_bgzf: null argument");
    if (strlen(chrom) > (size_t)kBzMaxName) return ctx_fail(ctx, MSD_EINVAL, "msd_per_base_bgzf: contig name longer than %d bytes", kBzMaxName);
    if ((rc = ensure_query_buffers(ctx))) return rc;
    const uint32_t tlen = ctx->tlen[tid];
    uint64_t n = 0;
    if ((rc = compute_runs(ctx, ctx->d_arena + ctx->depth_off[tid], tlen, (int32_t)tlen, IdentityF(), &n))) return rc;   // K8: runs on the device and in pinned memory
    if (start) *start = ctx->h_run_start; if (stop) *stop = ctx->h_run_stop; if (depth) *depth = ctx->h_run_val;
    if (n_runs) *n_runs = (int64_t)n;
    *bgzf = nullptr; *n_bytes = 0; *members = nullptr; *n_members = 0;
    // a short contig is not worth K10's fixed cost (a code built from its lines, ~20 launches, four small read-backs): the caller gets the
    // runs only and formats them itself (MSD_BGZF_MIN_RUNS, default 50,000 lines; 0 = always on the device)
    if (n < env_u64("MSD_BGZF_MIN_RUNS", 50000)) return MSD_OK;
    // K10: the sequence lives in bedgz_host.cuh (the same code runs under the CPU emulation test); the CRC step is K1c
    rc = bz_run(ctx->bz, ctx->s_compute, chrom, ctx->d_run_start, ctx->d_run_stop, ctx->d_run_val, ctx->h_run_start, ctx->h_run_stop, ctx->h_run_val, n, kSMs * 8,
                [](cudaStream_t s, const uint8_t* d_text, const BlockDesc* d_desc, int nb, uint32_t* ticket, const uint32_t* expected, uint32_t* status, uint32_t* errors, uint32_t* crc_out) {
                    launch_crc(s, d_text, d_desc, nb, ticket, expected, status, errors, crc_out);
                });
    ctx->launches += ctx->bz.launches;
    if (rc) return ctx_fail(ctx, rc, "%s", ctx->bz.err);
    if (n == 0) return MSD_OK;
    *bgzf = ctx->bz.h_out; *n_bytes = (int64_t)ctx->bz.out_bytes; *members = ctx->bz.members.data(); *n_members = (int64_t)ctx->bz.members.size();
    return MSD_OK;
}

// gen_quantized for one contig: K9 runs, then the reference's filter (bin != -1 and bin < len(lookup) = n_q-1, :136,:140) compacts the
// pinned arrays in place; *kept = the runs the reference yields
static int quantized_runs_impl(msd_ctx* ctx, int tid, const int64_t* quants, int n_q, const char* who, uint64_t* kept) {
    int rc = check_tid(ctx, tid, who); if (rc) return rc;
    if (!quants || n_q < 2 || n_q > kMaxQuants) return ctx_fail(ctx, MSD_EINVAL, "%s: need 2..%d quantize values", who, kMaxQuants);
    if ((rc = ensure_query_buffers(ctx))) return rc;
    const uint32_t tlen = ctx->tlen[tid];
    QuantF f; f.qa.n = n_q; for (int i = 0; i < n_q; i++) f.qa.q[i] = quants[i];
    // gen_quantized scans pos in 0 ..< arr.high-1 = tlen-1 (:132) but seeds from arr[0] even when that range is empty (:129)
    const uint32_t n_eff = tlen >= 2 ? tlen - 1 : (tlen ? 1u : 0u);
    // the reference only yields bins that have a label: bin != -1 and bin < len(lookup) = n_q-1 (:136,:140); the compaction keeps the order
    uint64_t w = 0;
    if ((rc = compute_runs(ctx, ctx->d_arena + ctx->depth_off[tid], n_eff, (int32_t)tlen, f, &w, n_q - 1))) return rc;
    *kept = w;
    return MSD_OK;
}

int msd_quantized_runs(msd_ctx* ctx, int tid, const int64_t* quants, int n_q, const int32_t** start, const int32_t** stop,
                       const int32_t** bin, int64_t* n_runs) {
    uint64_t w = 0;
    int rc = quantized_runs_impl(ctx, tid, quants, n_q, "msd_quantized_runs", &w); if (rc) return rc;
    if (start) *start = ctx->h_run_start; if (stop) *stop = ctx->h_run_stop; if (bin) *bin = ctx->h_run_val;
    if (n_runs) *n_runs = (int64_t)w;
    return MSD_OK;
}

int msd_quantized_bgzf(msd_ctx* ctx, int tid, const int64_t* quants, int n_q, const char* chrom, const char* const* labels, int n_labels, const uint8_t** bgzf, int64_t* n_bytes,
                       const msd_bgzf_member** members, int64_t* n_members, const int32_t** start, const int32_t** stop, const int32_t** bin, int64_t* n_runs) {
    if (!ctx) return MSD_EINVAL;
    if (!chrom || !labels || !bgzf || !n_bytes || !members || !n_members) return ctx_fail(ctx, MSD_EINVAL, "msd_quantized_bgzf: null argument");
    if (n_labels < n_q - 1) return ctx_fail(ctx, MSD_EINVAL, "msd_quantized_bgzf: %d labels for %d quantize values", n_labels, n_q);
    uint64_t w = 0;
    int rc = quantized_runs_impl(ctx, tid, quants, n_q, "msd_quantized_bgzf", &w); if (rc) return rc;
    if (start) *start = ctx->h_run_start; if (stop) *stop = ctx->h_run_stop; if (bin) *bin = ctx->h_run_val;
    if (n_runs) *n_runs = (int64_t)w;
    *bgzf = nullptr; *n_bytes = 0; *members = nullptr; *n_members = 0;
    if (w == 0 || w < env_u64("MSD_BGZF_MIN_RUNS", 50000)) return MSD_OK;      // short contig: runs only, see msd_per_base_bgzf
    cudaStream_t sc = ctx->s_compute;     // K9 left the kept runs on the device (and in pinned memory)
    rc = bz_run(ctx->bz, sc, chrom, ctx->d_run_start, ctx->d_run_stop, ctx->d_run_val, ctx->h_run_start, ctx->h_run_stop, ctx->h_run_val, w, kSMs * 8,
                [](cudaStream_t s, const uint8_t* d_text, const BlockDesc* d_desc, int nb, uint32_t* ticket, const uint32_t* expected, uint32_t* status, uint32_t* errors, uint32_t* crc_out) {
                    launch_crc(s, d_text, d_desc, nb, ticket, expected, status, errors, crc_out);
                }, labels, n_labels);
    ctx->launches += ctx->bz.launches;
    if (rc) return ctx_fail(ctx, rc, "msd_quantized_bgzf: %s", ctx->bz.err);
    *bgzf = ctx->bz.h_out; *n_bytes = (int64_t)ctx->bz.out_bytes; *members = ctx->bz.members.data(); *n_members = (int64_t)ctx->bz.members.size();
    return MSD_OK;
}

int msd_totals_device(msd_ctx* ctx, void** global_dist, void** region_dist, int64_t* n_bins, void** stats8) {
    if (!ctx) return MSD_EINVAL;
    cudaSetDevice(ctx->device);
    int rc = ensure_query_buffers(ctx); if (rc) return rc;
    if (global_dist) *global_dist = ctx->d_tot_global; if (region_dist) *region_dist = ctx->d_tot_region;
    if (n_bins) *n_bins = kDistBins; if (stats8) *stats8 = ctx->d_tot_stats;
    return MSD_OK;
}

int msd_totals_reset(msd_ctx* ctx) {
    if (!ctx) return MSD_EINVAL;
    cudaSetDevice(ctx->device);
    if (!ctx->d_tot_global) return ensure_query_buffers(ctx);      // allocates the accumulators and comes back here to zero them
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_tot_global, 0, 8ull * kDistBins, ctx->s_compute));
    MSD_CUDA_OK(cudaMemsetAsync(ctx->d_tot_region, 0, 8ull * kDistBins, ctx->s_compute));
    init_stats_kernel<<<1, 1, 0, ctx->s_compute>>>(ctx->d_tot_stats);
    MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
    return MSD_OK;
}

int msd_mark(msd_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) return MSD_EINVAL;
    cudaSetDevice(ctx->device);
    if (!ctx->marks[slot]) MSD_CUDA_OK(cudaEventCreate(&ctx->marks[slot]));
    MSD_CUDA_OK(cudaEventRecord(ctx->marks[slot], ctx->s_compute));
    return MSD_OK;
}

int msd_elapsed_ms(msd_ctx* ctx, int a, int b, double* ms) {
    if (!ctx || a < 0 || a >= 8 || b < 0 || b >= 8 || !ctx->marks[a] || !ctx->marks[b] || !ms) return MSD_EINVAL;
    cudaSetDevice(ctx->device);
    MSD_CUDA_OK(cudaEventSynchronize(ctx->marks[b]));
    float f = 0; MSD_CUDA_OK(cudaEventElapsedTime(&f, ctx->marks[a], ctx->marks[b]));
    *ms = f;
    return MSD_OK;
}

uint64_t msd_launch_count(const msd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t msd_debug_inflate(msd_ctx* ctx, const void* bgzf_bytes, uint64_t n_bytes, void* out, uint64_t out_cap) {
    if (!ctx || !bgzf_bytes || !out) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_inflate: bad arguments");
    cudaSetDevice(ctx->device);
    const uint8_t* F = (const uint8_t*)bgzf_bytes;
    std::vector<BlockDesc> desc; uint64_t pos = 0, infl = 0;
    while (pos < n_bytes) {
        uint32_t xlen = 0, isize = 0; uint32_t blen = bgzf_header(F, n_bytes, pos, &xlen, &isize);
        if (!blen) return ctx_fail(ctx, MSD_EIO, "invalid BGZF block header at %llu", (unsigned long long)pos);
        if (isize) { BlockDesc d; d.src = pos + 12 + xlen; d.clen = blen - xlen - 20; d.isize = isize; d.dst = infl; desc.push_back(d); infl += isize; }
        pos += blen;
    }
    if (infl > out_cap) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_inflate: need %llu bytes", (unsigned long long)infl);
    if (desc.empty()) return 0;
    uint8_t *d_c = nullptr, *d_o = nullptr; BlockDesc* d_d = nullptr; uint32_t *d_s = nullptr, *d_t = nullptr, *d_tok = nullptr, *d_nt = nullptr, *d_cr = nullptr; unsigned long long* d_rg = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &d_c, n_bytes + 256)) || (rc = dev_alloc(ctx, &d_o, infl + 256)) || (rc = dev_alloc(ctx, &d_d, desc.size())) ||
        (rc = dev_alloc(ctx, &d_s, desc.size())) || (rc = dev_alloc(ctx, &d_t, 8)) || (rc = dev_alloc(ctx, &d_tok, desc.size() * (uint64_t)kTokenCap)) ||
        (rc = dev_alloc(ctx, &d_nt, desc.size())) || (rc = dev_alloc(ctx, &d_cr, desc.size())) || (rc = dev_alloc(ctx, &d_rg, desc.size()))) { dev_free(&d_c); dev_free(&d_o); dev_free(&d_d); dev_free(&d_s); dev_free(&d_t); dev_free(&d_tok); dev_free(&d_nt); dev_free(&d_cr); dev_free(&d_rg); return rc; }
    cudaStream_t sc = ctx->s_compute;
    cudaMemcpyAsync(d_c, F, n_bytes, cudaMemcpyHostToDevice, sc);
    cudaMemcpyAsync(d_d, desc.data(), sizeof(BlockDesc) * desc.size(), cudaMemcpyHostToDevice, sc);
    cudaMemsetAsync(d_t, 0, 32, sc);
    const int nb = (int)desc.size();
    launch_inflate(sc, d_c, d_o, d_d, nb, d_t, d_tok, d_nt, d_s, d_t + 4, d_cr, !getenv("MSD_NO_CRC"), nullptr, nullptr, nullptr, d_rg);   // tickets: d_t[0..3]; error count: d_t[4]
    cudaError_t e = cudaGetLastError();
    uint32_t errs[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(errs, d_t + 3, 8, cudaMemcpyDeviceToHost, sc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, infl, cudaMemcpyDeviceToHost, sc);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sc);
    std::vector<uint32_t> status(desc.size());
    if (e == cudaSuccess && errs[1]) cudaMemcpy(status.data(), d_s, 4 * desc.size(), cudaMemcpyDeviceToHost);
    dev_free(&d_c); dev_free(&d_o); dev_free(&d_d); dev_free(&d_s); dev_free(&d_t); dev_free(&d_tok); dev_free(&d_nt); dev_free(&d_cr); dev_free(&d_rg);
    if (e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "msd_debug_inflate: %s", cudaGetErrorString(e));
    if (errs[1]) { size_t b = 0; while (b < status.size() && !status[b]) b++; return ctx_fail(ctx, MSD_EDATA, "%u block(s) failed to inflate; first: block %zu status %u", errs[1], b, b < status.size() ? status[b] : 0u); }
    return (int64_t)infl;
}

int64_t msd_debug_crc32(msd_ctx* ctx, const void* bytes, uint64_t n_bytes, uint32_t block_bytes, uint32_t first_offset, int reps,
                        uint32_t* crc_out, double* ms_per_launch) {
    if (!ctx || !bytes || !crc_out || n_bytes == 0 || block_bytes == 0 || block_bytes > 65536 || first_offset > 255 || reps < 0)
        return ctx_fail(ctx, MSD_EINVAL, "msd_debug_crc32: bad arguments");
    cudaSetDevice(ctx->device);
    const uint64_t nb64 = (n_bytes + block_bytes - 1) / block_bytes;
    if (nb64 > (1u << 30)) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_crc32: too many blocks");
    const int nb = (int)nb64;
    std::vector<BlockDesc> desc((size_t)nb);
    for (int i = 0; i < nb; i++) {
        desc[i].src = 0; desc[i].clen = 0; desc[i].dst = first_offset + (uint64_t)i * block_bytes;
        desc[i].isize = (uint32_t)std::min<uint64_t>(block_bytes, n_bytes - (uint64_t)i * block_bytes);
    }
    uint8_t* d_o = nullptr; BlockDesc* d_d = nullptr; uint32_t *d_s = nullptr, *d_t = nullptr, *d_exp = nullptr, *d_got = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &d_o, n_bytes + 512)) || (rc = dev_alloc(ctx, &d_d, (uint64_t)nb)) || (rc = dev_alloc(ctx, &d_s, (uint64_t)nb)) ||
        (rc = dev_alloc(ctx, &d_t, 4)) || (rc = dev_alloc(ctx, &d_exp, (uint64_t)nb)) || (rc = dev_alloc(ctx, &d_got, (uint64_t)nb))) {
        dev_free(&d_o); dev_free(&d_d); dev_free(&d_s); dev_free(&d_t); dev_free(&d_exp); dev_free(&d_got); return rc;
    }
    cudaStream_t sc = ctx->s_compute;
    cudaEvent_t e0 = nullptr, e1 = nullptr; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemsetAsync(d_o, 0, n_bytes + 512, sc);
    cudaMemcpyAsync(d_o + first_offset, bytes, n_bytes, cudaMemcpyHostToDevice, sc);
    cudaMemcpyAsync(d_d, desc.data(), sizeof(BlockDesc) * (size_t)nb, cudaMemcpyHostToDevice, sc);
    cudaMemsetAsync(d_s, 0, 4ull * nb, sc); cudaMemsetAsync(d_exp, 0, 4ull * nb, sc);
    for (int r = 0; r <= reps; r++) {
        cudaMemsetAsync(d_t, 0, 16, sc);
        if (r == 1) cudaEventRecord(e0, sc);
        launch_crc(sc, d_o, d_d, nb, d_t, d_exp, d_s, d_t + 1, d_got);
        if (r == 0) {   // the first launch marks every block whose CRC is not 0: from here on the expected values are the computed ones
            cudaMemcpyAsync(d_exp, d_got, 4ull * nb, cudaMemcpyDeviceToDevice, sc);
            cudaMemsetAsync(d_s, 0, 4ull * nb, sc);
        }
    }
    cudaEventRecord(e1, sc);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(crc_out, d_got, 4ull * nb, cudaMemcpyDeviceToHost, sc);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sc);
    float ms = 0; if (e == cudaSuccess && reps > 0) cudaEventElapsedTime(&ms, e0, e1);
    if (ms_per_launch) *ms_per_launch = reps > 0 ? (double)ms / reps : 0.0;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dev_free(&d_o); dev_free(&d_d); dev_free(&d_s); dev_free(&d_t); dev_free(&d_exp); dev_free(&d_got);
    if (e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "msd_debug_crc32: %s", cudaGetErrorString(e));
    return nb;
}

int msd_debug_depth(msd_ctx* ctx, int tid, int32_t* out, int64_t out_cap) {
    if (!ctx || !ctx->ran || tid < 0 || tid >= ctx->n_targets || !ctx->want[tid] || !out) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_depth: bad arguments");
    cudaSetDevice(ctx->device);
    const int64_t n = (int64_t)ctx->tlen[tid] + 1;
    if (out_cap < n) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_depth: need %lld cells", (long long)n);
    MSD_CUDA_OK(cudaMemcpy(out, ctx->d_arena + ctx->depth_off[tid], (size_t)n * 4, cudaMemcpyDeviceToHost));
    return MSD_OK;
}

int msd_debug_cumsum(msd_ctx* ctx, int32_t* inout, int64_t n) {
    if (!ctx || !inout || n < 0 || n > 0xffff0000ll) return ctx_fail(ctx, MSD_EINVAL, "msd_debug_cumsum: bad arguments");
    cudaSetDevice(ctx->device);
    int rc = ensure_query_buffers(ctx); if (rc) return rc;
    if (n == 0) return MSD_OK;
    int32_t* d = nullptr;
    if ((rc = dev_alloc(ctx, &d, (uint64_t)n + 32))) return rc;
    cudaError_t e = cudaMemcpyAsync(d, inout, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->s_compute);
    if (e == cudaSuccess) { std::vector<ScanSeg> one(1); one[0].arr_off = 0; one[0].n_scan = (uint32_t)n; one[0].n_stat = 0; rc = run_segments(ctx, d, one, true); if (rc) { dev_free(&d); return rc; } }
    if (e == cudaSuccess) e = cudaMemcpyAsync(inout, d, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    dev_free(&d);
    if (e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "msd_debug_cumsum: %s", cudaGetErrorString(e));
    return MSD_OK;
}

// ---- multi-GPU: spill exchange + totals all-reduce (comm.cuh) -----------------------------------------------------------------------
#define MSD_NCCL_OK(expr)                                                                                          \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess) return ctx_fail(ctx, MSD_ECOMM, "%s: %s", #expr, nccl_api()->GetErrorString(_r)); \
    } while (0)

int msd_comm_unique_id(void* id_out128) {
    if (!id_out128) return MSD_EINVAL;
    NcclApi* api = nccl_api();
    if (!api->handle) return ctx_fail(nullptr, MSD_ECOMM, "%s", api->why);
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return ctx_fail(nullptr, MSD_ECOMM, "ncclGetUniqueId: %s", api->GetErrorString(r));
    static_assert(sizeof id == MSD_COMM_ID_BYTES, "NCCL unique id size");
    memcpy(id_out128, &id, sizeof id);
    return MSD_OK;
}

int msd_comm_init(msd_ctx* ctx, int rank, int world, const void* id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return ctx_fail(ctx, MSD_EINVAL, "msd_comm_init: bad arguments");
    NcclApi* api = nccl_api();
    if (!api->handle) return ctx_fail(ctx, MSD_ECOMM, "%s", api->why);
    cudaSetDevice(ctx->device);
    msd_comm_destroy(ctx);
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    MSD_NCCL_OK(api->CommInitRank(&ctx->comm, world, id, rank));
    ctx->comm_rank = rank; ctx->comm_world = world;
    int rc;
    ctx->spill_words = (uint32_t)std::max<uint64_t>(env_u64("MSD_SPILL_CELLS", 1u << 20), 4096) + kSpillHeaderWords;
    if ((rc = dev_alloc(ctx, &ctx->d_spill_send, ctx->spill_words)) || (rc = dev_alloc(ctx, &ctx->d_spill_recv, ctx->spill_words)) ||
        (rc = dev_alloc(ctx, &ctx->d_comm_err, 4)) || (rc = dev_alloc(ctx, &ctx->d_stats_pack, 8))) return rc;
    MSD_CUDA_OK(cudaMallocHost((void**)&ctx->h_spill_hdr, 4 * kSpillHeaderWords));
    MSD_CUDA_OK(cudaMemset(ctx->d_comm_err, 0, 16));
    return MSD_OK;
}

int msd_comm_destroy(msd_ctx* ctx) {
    if (!ctx) return MSD_EINVAL;
    if (ctx->comm) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->s_compute); nccl_api()->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    dev_free(&ctx->d_spill_send); dev_free(&ctx->d_spill_recv); dev_free(&ctx->d_comm_err); dev_free(&ctx->d_stats_pack);
    if (ctx->h_spill_hdr) { cudaFreeHost(ctx->h_spill_hdr); ctx->h_spill_hdr = nullptr; }
    ctx->comm_world = 1; ctx->comm_rank = 0;
    return MSD_OK;
}

int msd_exchange_spills(msd_ctx* ctx) {
    if (!ctx || !ctx->comm) return ctx_fail(ctx, MSD_EINVAL, "msd_exchange_spills: call msd_comm_init first");
    if (!ctx->ran) return ctx_fail(ctx, MSD_EINVAL, "msd_exchange_spills: call msd_run first");
    cudaSetDevice(ctx->device);
    NcclApi* api = nccl_api();
    cudaStream_t sc = ctx->s_compute;
    const int rank = ctx->comm_rank, world = ctx->comm_world;
    // what this rank's reads leave beyond the end of every range it owns (max_end came back with msd_run)
    uint32_t* H = ctx->h_spill_hdr; uint32_t n_ent = 0; uint64_t cells = 0; std::vector<int> shards;
    for (int t = 0; t < ctx->n_targets; t++) {
        if (!ctx->want[t] || !ctx->ranged[t] || !ctx->pending[t]) continue;
        shards.push_back(t);
        const uint32_t tlen = ctx->tlen[t], hi = std::min(ctx->range_hi[t], tlen);
        if (hi >= tlen || ctx->max_end[t] <= hi) continue;
        const uint32_t n = std::min(ctx->max_end[t], tlen) - hi;
        if (n_ent >= kSpillMaxEntries) return ctx_fail(ctx, MSD_ENOMEM, "msd_exchange_spills: more than %u cut contigs on one rank", kSpillMaxEntries);
        H[1 + 3 * n_ent] = (uint32_t)t; H[2 + 3 * n_ent] = hi; H[3 + 3 * n_ent] = n; n_ent++; cells += n;
    }
    H[0] = n_ent;
    if (kSpillHeaderWords + cells > ctx->spill_words) return ctx_fail(ctx, MSD_ENOMEM, "msd_exchange_spills: %llu depth cells reach beyond this rank's ranges: raise MSD_SPILL_CELLS (on every rank)", (unsigned long long)cells);
    if (rank + 1 < world) {
        MSD_CUDA_OK(cudaMemcpyAsync(ctx->d_spill_send, H, 4ull * (1 + 3 * n_ent), cudaMemcpyHostToDevice, sc));
        if (n_ent) { spill_pack_kernel<<<kSMs, 256, 0, sc>>>(ctx->d_arena, (const unsigned long long*)ctx->d_depth_off, ctx->d_spill_send); ctx->launches++; }
    } else if (n_ent) return ctx_fail(ctx, MSD_EINVAL, "msd_exchange_spills: the last rank owns a range that ends inside a contig");
    MSD_NCCL_OK(api->GroupStart());
    if (rank + 1 < world) MSD_NCCL_OK(api->Send(ctx->d_spill_send, ctx->spill_words, ncclUint32, rank + 1, ctx->comm, sc));
    if (rank > 0) MSD_NCCL_OK(api->Recv(ctx->d_spill_recv, ctx->spill_words, ncclUint32, rank - 1, ctx->comm, sc));
    MSD_NCCL_OK(api->GroupEnd());
    if (rank > 0) {
        spill_unpack_kernel<<<kSMs, 256, 0, sc>>>(ctx->d_arena, (const unsigned long long*)ctx->d_depth_off, ctx->d_tlen, ctx->d_range_lo, ctx->d_range_hi, ctx->n_targets,
                                                  ctx->d_spill_recv, ctx->spill_words, ctx->d_comm_err);
        ctx->launches++;
    }
    MSD_CUDA_OK(cudaGetLastError());
    int rc = finalize_shards(ctx, shards); if (rc) return rc;
    uint32_t err = 0;
    MSD_CUDA_OK(cudaMemcpy(&err, ctx->d_comm_err, 4, cudaMemcpyDeviceToHost));
    if (err) { cudaMemset(ctx->d_comm_err, 0, 4); return ctx_fail(ctx, MSD_ESPAN, err == 2 ? "msd_exchange_spills: the previous rank's reads reach beyond this rank's whole range of a contig (shards too small)" : "msd_exchange_spills: the previous rank handed over depth for a range this rank does not own (shards of the ranks do not meet)"); }
    return MSD_OK;
}

int msd_totals_allreduce(msd_ctx* ctx) {
    if (!ctx || !ctx->comm) return ctx_fail(ctx, MSD_EINVAL, "msd_totals_allreduce: call msd_comm_init first");
    cudaSetDevice(ctx->device);
    int rc = ensure_query_buffers(ctx); if (rc) return rc;
    NcclApi* api = nccl_api();
    cudaStream_t sc = ctx->s_compute;
    stats_pack_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats, ctx->d_stats_pack);
    MSD_NCCL_OK(api->GroupStart());
    MSD_NCCL_OK(api->AllReduce(ctx->d_tot_global, ctx->d_tot_global, (size_t)kDistBins, ncclInt64, ncclSum, ctx->comm, sc));
    MSD_NCCL_OK(api->AllReduce(ctx->d_tot_region, ctx->d_tot_region, (size_t)kDistBins, ncclInt64, ncclSum, ctx->comm, sc));
    MSD_NCCL_OK(api->AllReduce(ctx->d_stats_pack, ctx->d_stats_pack, 4, ncclInt64, ncclSum, ctx->comm, sc));
    MSD_NCCL_OK(api->AllReduce(ctx->d_stats_pack + 4, ctx->d_stats_pack + 4, 2, ncclInt64, ncclMin, ctx->comm, sc));
    MSD_NCCL_OK(api->AllReduce(ctx->d_stats_pack + 6, ctx->d_stats_pack + 6, 2, ncclInt64, ncclMax, ctx->comm, sc));
    MSD_NCCL_OK(api->GroupEnd());
    stats_unpack_kernel<<<1, 1, 0, sc>>>(ctx->d_tot_stats, ctx->d_stats_pack);
    ctx->launches += 2;
    MSD_CUDA_OK(cudaGetLastError());
    return MSD_OK;
}

int msd_totals(msd_ctx* ctx, msd_depth_stat* global_stat, msd_depth_stat* region_stat, int64_t* global_dist, int64_t global_cap, int64_t* global_len,
               int64_t* region_dist, int64_t region_cap, int64_t* region_len) {
    if (!ctx) return MSD_EINVAL;
    cudaSetDevice(ctx->device);
    int rc = ensure_query_buffers(ctx); if (rc) return rc;
    long long st[8];
    MSD_CUDA_OK(cudaMemcpyAsync(st, ctx->d_tot_stats, sizeof st, cudaMemcpyDeviceToHost, ctx->s_compute));
    MSD_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
    auto put = [](msd_depth_stat* o, const long long* v) { if (!o) return; o->cum_length = v[0]; o->cum_depth = (uint64_t)v[1]; o->min_depth = (uint32_t)v[2]; o->max_depth = (uint32_t)v[3]; };
    put(global_stat, st); put(region_stat, st + 4);
    auto fetch = [&](const long long* d_bins, long long mx, int64_t floor_bins, int64_t* out, int64_t cap, int64_t* len, const char* what) -> int {
        if (!out && !len) return 0;
        const int64_t top = std::max<int64_t>(floor_bins - 1, (mx > 0x7fffffffll || mx < 0) ? (int64_t)kMaxCoverage : std::min<int64_t>(mx, kMaxCoverage));
        std::vector<int64_t> tmp((size_t)top + 1, 0);
        MSD_CUDA_OK(cudaMemcpy(tmp.data(), d_bins, 8ull * (top + 1), cudaMemcpyDeviceToHost));
        while (tmp.size() > 1 && tmp.back() == 0) tmp.pop_back();
        if (len) *len = (int64_t)tmp.size();
        if (out) { if (cap < (int64_t)tmp.size()) return ctx_fail(ctx, MSD_EINVAL, "msd_totals: %s capacity %lld < %zu bins", what, (long long)cap, tmp.size()); memcpy(out, tmp.data(), 8 * tmp.size()); }
        return 0;
    };
    if ((rc = fetch(ctx->d_tot_global, st[3], 1, global_dist, global_cap, global_len, "global_dist"))) return rc;
    if ((rc = fetch(ctx->d_tot_region, st[7], kWindowDistBins, region_dist, region_cap, region_len, "region_dist"))) return rc;   // window mode bins by mean: up to 511 whatever the depths
    return MSD_OK;
}

}  // extern "C"
