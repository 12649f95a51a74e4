## mosdepth_b200.nim -- {.importc, dynlib.} binding of libmosdepth_b200.so for the reference's Nim host.
## SOURCE ONLY: there is no Nim toolchain in the build image, so this file has not been compiled (INTEGRATION.md).
## It declares EVERY export of include/mosdepth_b200.h (ABI version 2; tests/test_abi_and_host.py compares the two lists); `check`
## turns a negative return into the reference's own failure mode (`quit msg`, mosdepth.nim:190-191).

const libmsd = "libmosdepth_b200.so"

type
  MsdCtx* = pointer
  MsdDepthStat* {.bycopy.} = object   # depth_stat, depthstat.nim:2-6
    cum_length*: int64
    cum_depth*: uint64
    min_depth*: uint32
    max_depth*: uint32

const
  MSD_MODE_DEFAULT* = 0.cint
  MSD_MODE_FAST* = 1.cint
  MSD_MODE_FRAGMENT* = 2.cint

type MsdRunInfo* {.bycopy.} = object   ## msd_run_info: device times (ms) and counts of the last msd_run
  ms_total*, ms_h2d*, ms_inflate*, ms_frame*, ms_records*, ms_scan*, ms_zero*: float64
  bytes_compressed*, bytes_inflated*, n_blocks*, n_records*, n_events*, n_depth_cells*: uint64
  n_launches*, n_inflate_launches*: uint32
  ms_inflate_span*: float64

proc msd_abi_version*(): cint {.importc, dynlib: libmsd.}
proc msd_create*(device: cint, ctx: ptr MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_destroy*(ctx: MsdCtx) {.importc, dynlib: libmsd.}
proc msd_last_error*(ctx: MsdCtx): cstring {.importc, dynlib: libmsd.}
proc msd_open*(ctx: MsdCtx, bam_path: cstring, n_targets: cint, target_len: ptr uint32, first_record_voff: uint64): cint {.importc, dynlib: libmsd.}
proc msd_open_mem*(ctx: MsdCtx, bam_bytes: pointer, n_bytes: uint64, n_targets: cint, target_len: ptr uint32, first_record_voff: uint64): cint {.importc, dynlib: libmsd.}
proc msd_stage*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_set_spans*(ctx: MsdCtx, n: cint, voff_beg, voff_end: ptr uint64): cint {.importc, dynlib: libmsd.}
proc msd_set_targets*(ctx: MsdCtx, want: ptr uint8): cint {.importc, dynlib: libmsd.}
proc msd_set_filters*(ctx: MsdCtx, mapq: cint, min_len, max_len: int64, eflag, iflag: uint16,
                      read_groups: cstringArray, n_read_groups: cint, mode: cint): cint {.importc, dynlib: libmsd.}
proc msd_run*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_last_run_info*(ctx: MsdCtx, info: ptr MsdRunInfo): cint {.importc, dynlib: libmsd.}
proc msd_contig_records*(ctx: MsdCtx, tid: cint, n: ptr int64): cint {.importc, dynlib: libmsd.}
proc msd_contig_stats*(ctx: MsdCtx, tid: cint, stat: ptr MsdDepthStat, dist: ptr int64, dist_cap: int64, dist_len: ptr int64): cint {.importc, dynlib: libmsd.}
proc msd_n_windows*(tlen, window: uint32): int64 {.importc, dynlib: libmsd.}
proc msd_prefetch_windows*(ctx: MsdCtx, window: uint32): cint {.importc, dynlib: libmsd.}
proc msd_set_contig_range*(ctx: MsdCtx, tid: cint, lo, hi: uint32): cint {.importc, dynlib: libmsd.}
proc msd_set_contig_cover*(ctx: MsdCtx, tid: cint, covered_from: uint32): cint {.importc, dynlib: libmsd.}
proc msd_contig_span_need*(ctx: MsdCtx, tid: cint, need_from, covered_from: ptr uint32): cint {.importc, dynlib: libmsd.}
proc msd_contig_spill*(ctx: MsdCtx, tid: cint, `from`, n: ptr uint32): cint {.importc, dynlib: libmsd.}
proc msd_depth_copy*(ctx: MsdCtx, tid: cint, pos, n: uint32, dst: ptr int32, dst_is_device: cint): cint {.importc, dynlib: libmsd.}
proc msd_depth_add*(ctx: MsdCtx, tid: cint, pos, n: uint32, src: ptr int32, src_is_device: cint): cint {.importc, dynlib: libmsd.}
proc msd_contig_finalize*(ctx: MsdCtx, tid: cint): cint {.importc, dynlib: libmsd.}
proc msd_windows*(ctx: MsdCtx, tid: cint, window: uint32, use_median: cint, value_out: ptr float64, value_cap: int64,
                  region_stat: ptr MsdDepthStat, region_dist512: ptr int64): int64 {.importc, dynlib: libmsd.}
proc msd_regions*(ctx: MsdCtx, tid: cint, n: int64, start, stop: ptr uint32, use_median: cint, value_out: ptr float64,
                  region_stat: ptr MsdDepthStat, region_dist: ptr int64, dist_cap: int64, dist_len: ptr int64,
                  thresholds: ptr int32, n_thr: cint, thr_counts: ptr int64): cint {.importc, dynlib: libmsd.}
proc msd_per_base_runs*(ctx: MsdCtx, tid: cint, start, stop, depth: ptr ptr int32, n_runs: ptr int64): cint {.importc, dynlib: libmsd.}
type MsdBgzfMember* {.bycopy.} = object   ## msd_bgzf_member
  offset*, first_run*: uint64
  csize*, isize*: uint32
## per-base.bed.gz members deflated on the device (replaces the loop :783-793 and the BGZF compression of its writer, :650-651)
proc msd_per_base_bgzf*(ctx: MsdCtx, tid: cint, chrom: cstring, bgzf: ptr ptr uint8, n_bytes: ptr int64, members: ptr ptr MsdBgzfMember, n_members: ptr int64,
                        start, stop, depth: ptr ptr int32, n_runs: ptr int64): cint {.importc, dynlib: libmsd.}
proc msd_quantized_bgzf*(ctx: MsdCtx, tid: cint, quants: ptr int64, n_q: cint, chrom: cstring, labels: cstringArray, n_labels: cint, bgzf: ptr ptr uint8, n_bytes: ptr int64,
                         members: ptr ptr MsdBgzfMember, n_members: ptr int64, start, stop, bin: ptr ptr int32, n_runs: ptr int64): cint {.importc, dynlib: libmsd.}
proc msd_quantized_runs*(ctx: MsdCtx, tid: cint, quants: ptr int64, n_q: cint, start, stop, bin: ptr ptr int32, n_runs: ptr int64): cint {.importc, dynlib: libmsd.}

## genome-wide accumulators (`total` rows) and the collectives of the path: one context per GPU and per rank (process or thread)
proc msd_totals_device*(ctx: MsdCtx, global_dist, region_dist: ptr pointer, n_bins: ptr int64, stats8: ptr pointer): cint {.importc, dynlib: libmsd.}
proc msd_totals_reset*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_comm_unique_id*(id_out128: pointer): cint {.importc, dynlib: libmsd.}
proc msd_comm_init*(ctx: MsdCtx, rank, world: cint, id128: pointer): cint {.importc, dynlib: libmsd.}
proc msd_comm_destroy*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_exchange_spills*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_totals_allreduce*(ctx: MsdCtx): cint {.importc, dynlib: libmsd.}
proc msd_totals*(ctx: MsdCtx, global_stat, region_stat: ptr MsdDepthStat, global_dist: ptr int64, global_cap: int64, global_len: ptr int64,
                 region_dist: ptr int64, region_cap: int64, region_len: ptr int64): cint {.importc, dynlib: libmsd.}
## measurement and test hooks
proc msd_mark*(ctx: MsdCtx, slot: cint): cint {.importc, dynlib: libmsd.}
proc msd_elapsed_ms*(ctx: MsdCtx, slot_from, slot_to: cint, ms: ptr float64): cint {.importc, dynlib: libmsd.}
proc msd_launch_count*(ctx: MsdCtx): uint64 {.importc, dynlib: libmsd.}
proc msd_debug_inflate*(ctx: MsdCtx, bgzf_bytes: pointer, n_bytes: uint64, dst: pointer, out_cap: uint64): int64 {.importc, dynlib: libmsd.}
proc msd_debug_depth*(ctx: MsdCtx, tid: cint, dst: ptr int32, out_cap: int64): cint {.importc, dynlib: libmsd.}
proc msd_debug_cumsum*(ctx: MsdCtx, inout: ptr int32, n: int64): cint {.importc, dynlib: libmsd.}
proc msd_debug_crc32*(ctx: MsdCtx, bytes: pointer, n_bytes: uint64, block_bytes, first_offset: uint32, reps: cint, crc_out: ptr uint32, ms_per_launch: ptr float64): int64 {.importc, dynlib: libmsd.}

template check*(ctx: MsdCtx, call: untyped) =
  ## negative return -> the reference's own failure mode for iterator errors (mosdepth.nim:190-191)
  if call < 0: quit $msd_last_error(ctx)

# ---------------------------------------------------------------------------------------------------------------
# How main() (mosdepth.nim:589-840) changes.  `coverage()` (:250-360) and `to_coverage()` (:54-56) are replaced
# by one msd_run() before the per-target loop; inside the loop every consumer of `arr` becomes one call:
#
#   var ctx: MsdCtx
#   if msd_create(0, addr ctx) != 0: quit $msd_last_error(nil)
#   var tlens = newSeq[uint32](targets.len); for t in targets: tlens[t.tid] = t.length
#   check ctx, msd_open(ctx, bam_path, targets.len.cint, addr tlens[0], first_record_voff)  # bgzf_tell after the header
#   check ctx, msd_set_spans(ctx, n, addr begs[0], addr ends[0])       # from bam.idx: per-contig chunk min/max
#   check ctx, msd_set_targets(ctx, addr want[0])                      # get_targets + skip rule :703-705
#   check ctx, msd_set_filters(ctx, mapq.cint, min_len, max_len, eflag, iflag, rgs, n_rg, mode)
#   check ctx, msd_run(ctx)
#   for target in sub_targets:
#     var nrec: int64
#     let tid = msd_contig_records(ctx, target.tid.cint, addr nrec)    # coverage()'s tid / -2 (:707-712)
#     if region != "" and tid != -2:
#       if region.isdigit: discard msd_windows(ctx, tid, window, use_median.cint, addr vals[0], vals.len, addr chrom_region_stat, addr d512[0])
#       else:              check ctx, msd_regions(ctx, tid, n, addr starts[0], addr stops[0], ...)   # :720-743
#     if tid != -2: check ctx, msd_contig_stats(ctx, tid, addr chrom_stat, addr dist[0], dist.len, addr dist_len)  # :746-747
#     if not skip_per_base and tid != -2: check ctx, msd_per_base_runs(ctx, tid, addr s, addr e, addr d, addr n)   # :785
#     if quantize.len != 0 and tid != -2: check ctx, msd_quantized_runs(ctx, tid, addr quantize[0], quantize.len.cint, ...)  # :802
#   # formatting (write_summary, write_distribution, fastIntToStr lines, BGZI writers) is unchanged.
