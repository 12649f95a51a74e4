// common.cuh -- shared declarations for libmosdepth_b200 (sm_100a only; no CPU fallback anywhere in csrc/).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define MSD_HD __host__ __device__ __forceinline__
#define MSD_D __device__ __forceinline__
#else
#define MSD_HD inline
#define MSD_D inline
#endif

namespace msd {

constexpr int kSMs = 148;                 // B200: 2 dies x 74 SMs
constexpr uint32_t kMaxCoverage = 400000; // MAX_COVERAGE, mosdepth.nim:415
constexpr int kDistBins = 400001;         // bins 0..400000 (values > 400000 land in 399990, :429-430)
constexpr int kWindowDistBins = 512;      // region_distribution in window mode, mosdepth.nim:630,738-739

// One BGZF block as the inflate kernel sees it (built on the host while chasing BSIZE).
struct BlockDesc {
    uint64_t src;    // byte offset of the raw DEFLATE stream inside the compressed chunk buffer
    uint32_t clen;   // DEFLATE bytes (BSIZE+1 - header - 8)
    uint32_t isize;  // ISIZE trailer field
    uint64_t dst;    // byte offset of the block's output inside the inflated chunk buffer
};

// BAM fixed-length core, offsets from the start of the record (the block_size field itself is at 0).
// [ext] SAM/BAM spec section 4.2; SURVEY.md Appendix B.
constexpr int kRecBlockSize = 0, kRecRefID = 4, kRecPos = 8, kRecLName = 12, kRecMapq = 13, kRecNCigar = 16,
              kRecFlag = 18, kRecLSeq = 20, kRecNextRefID = 24, kRecNextPos = 28, kRecTlen = 32, kRecQName = 36;

constexpr uint16_t kFlagProper = 2, kFlagUnmap = 4, kFlagRead2 = 128, kFlagSupp = 2048;

// unaligned little-endian loads from a byte buffer (BAM records are not 4-byte aligned)
MSD_D uint32_t ld_u32(const uint8_t* __restrict__ buf, uint64_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(buf + (off & ~3ull));
    uint32_t sh = (uint32_t)(off & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    uint32_t hi = w[1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (lo >> sh) | (hi << (32 - sh));
#endif
}
MSD_D uint16_t ld_u16(const uint8_t* __restrict__ buf, uint64_t off) { return (uint16_t)(buf[off] | (buf[off + 1] << 8)); }

// bam_cigar_type (htslib sam.h): bit0 = consumes query, bit1 = consumes reference
MSD_HD int cigar_type(uint32_t op) { return (0x3C1A7 >> ((op & 15) << 1)) & 3; }

#define MSD_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return ctx_fail(ctx, MSD_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

}  // namespace msd
