/*
 * mosdepth_b200.h -- C ABI of libmosdepth_b200.so: the depth hot path of brentp/mosdepth on one B200 (sm_100a).
 *
 * The reference (Nim, /root/reference/mosdepth.nim, v0.3.14) has no plugin/FFI seam for this path; every
 * step is a direct call inside one module.  The seam is therefore cut at the reference's own function
 * boundaries; each entry point below names the reference code it replaces.  The host
 * (the Nim CLI in the reference; the C++ stand-in `mosdepth-b200` here) keeps: flag parsing, BAM header and
 * BAI/CSI parsing, BED loading, and all text/BGZF output writing.
 *
 * Conventions: every call returns >= 0 on success and a negative MSD_E* code on failure, message via
 * msd_last_error().  One host thread per context.  There is NO CPU fallback: msd_create() fails unless an
 * sm_100 device is present.  All pointers are plain host pointers; sizes are in elements unless noted.
 */
#ifndef MOSDEPTH_B200_H
#define MOSDEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSD_ABI_VERSION 2

/* error codes */
#define MSD_OK 0
#define MSD_ENODEV (-10)     /* no sm_100 device / CUDA failure at create */
#define MSD_EINVAL (-11)     /* bad argument or call order */
#define MSD_EIO (-12)        /* file / BGZF framing error */
#define MSD_ECUDA (-13)      /* CUDA runtime error */
#define MSD_EDATA (-14)      /* corrupt DEFLATE stream or BAM record (reference: quit on hts.BamError, mosdepth.nim:190-191) */
#define MSD_ENOMEM (-15)     /* a device-side table overflowed (message says which knob) */
#define MSD_ENEGDEPTH (-16)  /* negative depth met in median mode (reference raises IndexDefect, depthstat.nim:18-19) */
#define MSD_ESPAN (-17)      /* a shard's span does not hold every record its owned records depend on (msd_set_contig_range) */
#define MSD_ECOMM (-18)      /* NCCL missing or a collective failed */

/* msd_set_filters mode: which branch of coverage() is taken (mosdepth.nim:292,342,346,355) */
#define MSD_MODE_DEFAULT 0   /* CIGAR blocks + mate-overlap correction */
#define MSD_MODE_FAST 1      /* -x / --fast-mode */
#define MSD_MODE_FRAGMENT 2  /* -a / --fragment-mode */

/* mirrors depth_stat, depthstat.nim:2-6 */
typedef struct msd_depth_stat {
    int64_t cum_length;
    uint64_t cum_depth;
    uint32_t min_depth;   /* UINT32_MAX when nothing was accumulated (clear(), depthstat.nim:47-51) */
    uint32_t max_depth;
} msd_depth_stat;

/* per-stage device times of the last msd_run(), CUDA events on the library's own stream (milliseconds) and
 * the byte/record counts a roofline needs.  Filled by msd_last_run_info(). */
typedef struct msd_run_info {
    double ms_total;           /* first launch .. last launch of msd_run on the compute stream */
    double ms_h2d;             /* summed host->device copies of compressed bytes (0 when staged) */
    double ms_inflate;         /* summed bgzf_inflate kernel time (launches overlap: see ms_inflate_span) */
    double ms_frame;           /* summed framing kernels */
    double ms_records;         /* summed filter / delta kernels (+ mate join in default mode) */
    double ms_scan;            /* summed prefix-sum + contig stats kernels */
    double ms_zero;            /* depth arena zero-fill */
    uint64_t bytes_compressed; /* BGZF bytes consumed (whole blocks) */
    uint64_t bytes_inflated;   /* bytes produced by inflate */
    uint64_t n_blocks;         /* BGZF blocks inflated */
    uint64_t n_records;        /* BAM records framed (before any filter) */
    uint64_t n_events;         /* +1/-1 events applied to the depth arrays */
    uint64_t n_depth_cells;    /* sum of (tlen+1) over contigs that had records */
    uint32_t n_launches;       /* kernels launched by the library during the run */
    uint32_t n_inflate_launches;
    double ms_inflate_span;    /* first inflate launch start .. last inflate launch end (launches of consecutive chunks overlap) */
} msd_run_info;

typedef struct msd_ctx msd_ctx;

/* -- lifetime ------------------------------------------------------------------------------------------- */
int msd_abi_version(void);
/* device: CUDA ordinal.  Fails with MSD_ENODEV unless the device is compute capability 10.x. */
int msd_create(int device, msd_ctx **out);
void msd_destroy(msd_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed msd_create on this thread. */
const char *msd_last_error(const msd_ctx *ctx);

/* -- input: replaces open(bam, threads, index=true) mosdepth.nim:968 and bam.hdr.targets :596 ----------- *
 * The host has parsed the BAM header: it passes the target lengths and the virtual offset
 * (coffset<<16 | uoffset) of the first alignment record.  The library reads the compressed bytes itself. */
int msd_open(msd_ctx *ctx, const char *bam_path, int n_targets, const uint32_t *target_len, uint64_t first_record_voff);
/* Same, for a BAM image already in host memory (pinned or pageable); the caller keeps it alive. */
int msd_open_mem(msd_ctx *ctx, const void *bam_bytes, uint64_t n_bytes, int n_targets, const uint32_t *target_len,
                 uint64_t first_record_voff);
/* Optional: copy the compressed bytes of the current spans into HBM now, so later msd_run() calls start from
 * device-resident input (no host->device traffic inside the run). */
int msd_stage(msd_ctx *ctx);

/* -- what to read: replaces the index iterator behind bam.query(tid, 0, len), mosdepth.nim:185 ----------- *
 * n virtual-offset ranges [beg, end) taken by the host from the BAI/CSI (pseudo-bin ref_beg/ref_end or bin
 * chunks); each beg must be the start of a record, end = 0 means "to end of file".  Default (never called, or
 * n = 0): one span from first_record_voff to end of file.  want_target (n_targets bytes, may be NULL = all)
 * selects the contigs whose records are counted: the per-contig loop's skip rule (mosdepth.nim:703-705), -c, and
 * the multi-GPU shard assignment all reduce to this mask. */
int msd_set_spans(msd_ctx *ctx, int n, const uint64_t *voff_beg, const uint64_t *voff_end);
int msd_set_targets(msd_ctx *ctx, const uint8_t *want_target);

/* -- coverage() parameters, mosdepth.nim:250-254 (filters :276-285) ---------------------------------------- */
int msd_set_filters(msd_ctx *ctx, int mapq, int64_t min_len, int64_t max_len, uint16_t eflag, uint16_t iflag,
                    const char *const *read_groups, int n_read_groups, int mode);

/* -- coverage() + to_coverage() for every selected contig, mosdepth.nim:707,713 --------------------------- *
 * BGZF inflate -> record framing -> filters -> +1/-1 events (-> mate-overlap correction) -> per-contig int32
 * prefix sum, fused with the whole-contig depth distribution (:746) and depth_stat (:747).  After it returns,
 * the depth arrays of all selected contigs are resident in HBM and the query calls below can be made in any
 * order. */
int msd_run(msd_ctx *ctx);
int msd_last_run_info(const msd_ctx *ctx, msd_run_info *out);

/* coverage()'s return convention (mosdepth.nim:255-258): returns tid when the contig had >= 1 record (counted
 * BEFORE the filters, :272-275), -2 when it had none, -1 when tid is out of range / not selected. */
int msd_contig_records(msd_ctx *ctx, int tid, int64_t *n_records_prefilter);

/* inc(chrom_global_distribution, arr, 0, len-1) :746 and newDepthStat(arr[0..<len-1]) :747.
 * dist: caller array of dist_cap int64 bins; *dist_len = 1 + highest non-zero bin (values > 400000 land in
 * bin 399990, negatives are skipped, :415-437).  MSD_EINVAL if dist_cap is too small (*dist_len is still set). */
int msd_contig_stats(msd_ctx *ctx, int tid, msd_depth_stat *chrom_stat, int64_t *dist, int64_t dist_cap, int64_t *dist_len);

/* window loop mosdepth.nim:720-741 with window_gen :382-388 evaluated on the device.
 * value_out[n_windows]: imean() result per window (:399-413): the sequential float64 mean, or the CountStat
 * median when use_median != 0.  region_stat: sum of newDepthStat over the windows (:723-724).
 * region_dist[512]: bin min(toInt(mean), 511) per window (:737-739).  Returns n_windows written;
 * value_cap is the capacity of value_out (MSD_EINVAL if too small; msd_n_windows() gives the count). */
int64_t msd_n_windows(uint32_t tlen, uint32_t window);
/* Optional hint, call before msd_run(): the host is going to ask msd_windows(tid, window, use_median = 0) for every
 * contig (mosdepth --by <window>, the loop :720-741 inside `for target in sub_targets` :691).  msd_run() then computes
 * the windows of all contigs in one launch and msd_windows() returns them from pinned host memory -- same values, no
 * launch/sync latency per contig; their contribution to the genome-wide `total_region` accumulators (msd_totals) is made by
 * msd_run() for every such contig, whether or not msd_windows() is called for it.  window = 0 switches the hint off.  Nothing
 * in the reference corresponds to it. */
int msd_prefetch_windows(msd_ctx *ctx, uint32_t window);
int64_t msd_windows(msd_ctx *ctx, int tid, uint32_t window, int use_median, double *value_out, int64_t value_cap,
                    msd_depth_stat *region_stat, int64_t *region_dist512);

/* BED loop mosdepth.nim:720-743 for n host-sorted regions of one contig.
 * value_out[n] as above; region_stat accumulates newDepthStat(arr[min(start,L)..<min(L,stop)]) with L = tlen+1
 * (:719,:724); region_dist: per-base histogram inside the regions (inc, :741), same convention as
 * msd_contig_stats; thresholds (ascending, n_thr may be 0): thr_counts[i*n_thr + j] = #bases of region i with
 * depth >= thresholds[j] (:549-553; slices are clamped to the array, see DESIGN.md "Divergences").
 * region_stat == NULL: the call does not count towards the genome-wide `total_region` accumulators (msd_totals) -- the host
 * uses that for the thresholds of --by <window>, whose windows msd_windows() has already accounted for. */
int msd_regions(msd_ctx *ctx, int tid, int64_t n, const uint32_t *start, const uint32_t *stop, int use_median,
                double *value_out, msd_depth_stat *region_stat, int64_t *region_dist, int64_t dist_cap,
                int64_t *dist_len, const int32_t *thresholds, int n_thr, int64_t *thr_counts);

/* gen_depths(arr) mosdepth.nim:58-96 as consumed at :785-793: run-length segments of arr[0..tlen).
 * The three arrays are library-owned pinned host memory, valid until the next call on this context that
 * returns runs.  stop[i] == start[i+1]; stop[n-1] == tlen. */
int msd_per_base_runs(msd_ctx *ctx, int tid, const int32_t **start, const int32_t **stop, const int32_t **depth,
                      int64_t *n_runs);

/* The per-base.bed.gz bytes of a contig, formatted and deflated on the device: replaces the loop mosdepth.nim:783-793
 * (`chrom \t start \t stop \t depth \n` per run of gen_depths, numbers as int2str.nim:59-76 prints them) together with
 * the BGZF compression hts-nim's writer does for it (wopen_bgzi :650-651).  *bgzf points at *n_bytes of complete BGZF
 * members (no EOF marker: the caller appends it when it closes the file) that decompress to exactly that text; the
 * compressed bytes differ from zlib's (parity is on decompressed content).  members[k] describes member k: its
 * byte offset in *bgzf, the index of the run its first line holds (a line never straddles two members), its compressed
 * and uncompressed sizes -- what the caller needs to build the .csi from the runs, which are returned as by
 * msd_per_base_runs().  All buffers are library-owned pinned host memory, valid until the next call on this context that
 * returns runs.  chrom: at most 200 bytes (MSD_EINVAL beyond; the caller then formats the runs itself).
 * A contig with fewer than MSD_BGZF_MIN_RUNS runs (environment, default 50,000) is not worth the fixed cost of the device path: the call
 * then returns the runs with *n_members = 0 and the caller formats them itself (as after msd_per_base_runs). */
typedef struct msd_bgzf_member {
    uint64_t offset;      /* of the member inside *bgzf */
    uint64_t first_run;   /* index of the run whose line opens the member */
    uint32_t csize;       /* member bytes (BSIZE + 1) */
    uint32_t isize;       /* text bytes */
} msd_bgzf_member;
int msd_per_base_bgzf(msd_ctx *ctx, int tid, const char *chrom, const uint8_t **bgzf, int64_t *n_bytes,
                      const msd_bgzf_member **members, int64_t *n_members, const int32_t **start,
                      const int32_t **stop, const int32_t **depth, int64_t *n_runs);

/* gen_quantized(quants, arr) mosdepth.nim:124-141 with linear_search :98-109.  quants: the n_q sorted values of
 * get_quantize_args (:501-519; INT64_MAX for an open last bin).  Returns only the segments the reference
 * yields (bin != -1 and bin < n_q-1; the scan stops before the last base, :132; the last segment is closed
 * at tlen, :140-141).  bin indexes the host's make_lookup() labels (:111-122). */
int msd_quantized_runs(msd_ctx *ctx, int tid, const int64_t *quants, int n_q, const int32_t **start,
                       const int32_t **stop, const int32_t **bin, int64_t *n_runs);

/* quantized.bed.gz the same way: msd_quantized_runs() followed by the formatting (`chrom \t start \t stop \t labels[bin] \n`,
 * mosdepth.nim:802-805) and BGZF compression of its lines on the device.  labels: the host's make_lookup() strings (:111-122),
 * n_labels >= n_q - 1, each at most 48 bytes (MSD_EINVAL beyond).  Outputs as for msd_per_base_bgzf(); the runs are those of
 * msd_quantized_runs(). */
int msd_quantized_bgzf(msd_ctx *ctx, int tid, const int64_t *quants, int n_q, const char *chrom, const char *const *labels,
                       int n_labels, const uint8_t **bgzf, int64_t *n_bytes, const msd_bgzf_member **members,
                       int64_t *n_members, const int32_t **start, const int32_t **stop, const int32_t **bin,
                       int64_t *n_runs);

/* -- multi-GPU: one context per GPU ------------------------------------------------------------------------ *
 * Device pointer + element count of this context's genome-wide accumulators, kept up to date by
 * msd_run / msd_contig_finalize (global) and msd_windows / msd_regions (region) -- sum_into, mosdepth.nim:761-764:
 * int64 global dist bins, int64 region dist bins, and {cum_length, cum_depth, min, max} x {global, region} packed
 * as 8 int64.  msd_run() zeroes them when it starts; msd_totals_reset() zeroes them on request. */
int msd_totals_device(msd_ctx *ctx, void **global_dist, void **region_dist, int64_t *n_bins, void **stats8);
int msd_totals_reset(msd_ctx *ctx);

/* The collectives of the path (north_star: "only the small dist/summary histograms are combined, via NCCL
 * allreduce over NVLink, and prefix-sum carries cross shard boundaries inside a contig").  Nothing in the single-process
 * reference corresponds to them.  One context per GPU and per rank; the ranks may be processes (bench.py under torchrun)
 * or host threads of one process (mosdepth-b200 --gpus N).  NCCL is loaded with dlopen("libnccl.so.2") on first use:
 * a process that never calls these needs no NCCL.
 *   msd_comm_unique_id : rank 0 obtains the 128-byte NCCL id and hands it to the other ranks by any means.
 *   msd_comm_init      : joins the communicator (collective: every rank calls it); rank r's shards are followed by
 *                        rank r+1's in genome order.
 *   msd_exchange_spills: after msd_run() on every rank.  For every contig this rank owns a range of
 *                        (msd_set_contig_range): the depth its reads leave beyond the range end goes to rank+1
 *                        (ncclSend straight out of the depth arena's staging buffer over NVLink), what rank-1 hands over is
 *                        added at the range start, and every shard is finalised (msd_contig_finalize): one fixed-size
 *                        message per neighbour, no host round trip between receive and add.  Replaces
 *                        msd_contig_spill + msd_depth_copy + msd_depth_add + msd_contig_finalize per shard.
 *   msd_totals_allreduce: sum_into / depth_stat `+` across ranks (mosdepth.nim:761-764, 807-813): ncclAllReduce of the two
 *                        dist histograms (sum) and the eight stats (sum / min / max) in place in every rank's accumulators.
 *   msd_totals         : read the accumulators back: the `total` and `total_region` rows.  dist arrays as in
 *                        msd_contig_stats (length = 1 + highest non-zero bin, at least 1). */
#define MSD_COMM_ID_BYTES 128
int msd_comm_unique_id(void *id_out128);
int msd_comm_init(msd_ctx *ctx, int rank, int world, const void *id128);
int msd_comm_destroy(msd_ctx *ctx);
int msd_exchange_spills(msd_ctx *ctx);
int msd_totals_allreduce(msd_ctx *ctx);
int msd_totals(msd_ctx *ctx, msd_depth_stat *global_stat, msd_depth_stat *region_stat, int64_t *global_dist, int64_t global_cap,
               int64_t *global_len, int64_t *region_dist, int64_t region_cap, int64_t *region_len);

/* ---- intra-contig sharding across GPUs (nothing in the single-process reference corresponds to it) --------
 * One context per GPU.  msd_set_contig_range(tid, lo, hi) before msd_run() makes this context the OWNER of the records of
 * contig tid with lo <= pos < hi (hi = 0: to the contig end; lo a multiple of 32, and of the window / 16384 if windows
 * will be asked for).  The spans given to msd_set_spans() must cover every owned record, at least one record at or
 * beyond hi (msd_run fails otherwise: extend the span) and, in default mode, the mates in front of lo that owned records
 * take from the `seen` table (they go through the filters and the table but emit nothing).  The caller declares how far back
 * its spans reach with msd_set_contig_cover(tid, covered_from): "every record of the contig with pos >= covered_from is in the
 * spans" (0: from the contig's first record; an index-derived span start that is the linear-index entry of the 16,384-bp
 * window w covers w * 16384).  msd_run checks it: an owned read that finds no mate in `seen` although its mate position lies
 * before covered_from fails the run with MSD_ESPAN (extend the span start; msd_contig_span_need() says how far) instead of
 * silently skipping an overlap correction.  Never declared: covered_from = lo.  After msd_run() the contig's array holds the prefix sum of this shard's own events.  Reads that start
 * before hi still cover [hi, hi + n): msd_contig_spill() says how far; that piece of depth is copied out
 * (msd_depth_copy) and added by the next shard (msd_depth_add) -- the prefix-sum carry across the shard boundary, as
 * depth rather than as one integer, because the reads of a shard all end within one read length of it.  Host or device
 * pointers (device: send/receive the piece with NCCL straight out of / into GPU memory).  msd_contig_finalize() then
 * computes the shard's part of msd_contig_stats() (cum_length = hi - lo; the caller sums the shards), after which
 * msd_windows() returns the windows that start in [lo, hi).  msd_regions / msd_per_base_runs / msd_quantized_runs are
 * not available on a shard; fragment mode is not supported with ranges (a fragment starts before its read 1).
 * msd_contig_records() counts owned records only. */
int msd_set_contig_range(msd_ctx *ctx, int tid, uint32_t lo, uint32_t hi);
int msd_set_contig_cover(msd_ctx *ctx, int tid, uint32_t covered_from);
/* After msd_run() failed with MSD_ESPAN: the position the span of contig tid has to reach back to (the smallest mate
 * position an owned read looked for in vain) and the covered_from the run was checked against. */
int msd_contig_span_need(msd_ctx *ctx, int tid, uint32_t *need_from, uint32_t *covered_from);
int msd_contig_spill(msd_ctx *ctx, int tid, uint32_t *from, uint32_t *n);
int msd_depth_copy(msd_ctx *ctx, int tid, uint32_t pos, uint32_t n, int32_t *dst, int dst_is_device);
int msd_depth_add(msd_ctx *ctx, int tid, uint32_t pos, uint32_t n, const int32_t *src, int src_is_device);
int msd_contig_finalize(msd_ctx *ctx, int tid);

/* -- measurement hooks (bench.py): CUDA events on the library's own compute stream --------------------- */
/* Record event `slot` (0..7) on the compute stream now. */
int msd_mark(msd_ctx *ctx, int slot);
/* Milliseconds between two recorded marks (synchronises on the later one). */
int msd_elapsed_ms(msd_ctx *ctx, int slot_from, int slot_to, double *ms);
/* Kernels launched by this context since it was created. */
uint64_t msd_launch_count(const msd_ctx *ctx);

/* -- test hooks (used by tests/ to check single stages against zlib / numpy) ----------------------------- */
/* Inflate every BGZF block of a host BGZF image; out must hold the sum of ISIZE fields. Returns bytes written. */
int64_t msd_debug_inflate(msd_ctx *ctx, const void *bgzf_bytes, uint64_t n_bytes, void *out, uint64_t out_cap);
/* Copy depth[0..tlen] (tlen+1 int32, after the prefix sum) of a contig to host memory. */
int msd_debug_depth(msd_ctx *ctx, int tid, int32_t *out, int64_t out_cap);
/* In-place inclusive int32 prefix sum of a host array on the device (the scan kernel alone). */
int msd_debug_cumsum(msd_ctx *ctx, int32_t *inout, int64_t n);
/* K1c alone: the host bytes are placed `first_offset` bytes past a 256-byte boundary in device memory and cut into blocks of
 * `block_bytes` (the last one shorter; block_bytes <= 65536); crc_out[i] receives the CRC-32 of block i (RFC 1952, what the
 * gzip trailer holds).  The kernel is launched 1 + reps times; *ms_per_launch (may be NULL) is the average of the last `reps`
 * launches, timed with CUDA events on the launching stream.  Returns the number of blocks. */
int64_t msd_debug_crc32(msd_ctx *ctx, const void *bytes, uint64_t n_bytes, uint32_t block_bytes, uint32_t first_offset, int reps,
                        uint32_t *crc_out, double *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* MOSDEPTH_B200_H */
