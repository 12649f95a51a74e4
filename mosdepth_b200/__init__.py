"""mosdepth_b200 -- ctypes binding of libmosdepth_b200.so (the C ABI in include/mosdepth_b200.h).

This Python layer is plumbing for tests and bench.py only: it loads the CUDA library, parses BAM headers and
BAI/CSI indexes on the host (what the reference's host keeps doing) and forwards to the C ABI.
There is no CPU fallback: if the library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import hostio  # noqa: F401

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["MSD_LIB_PATH"]) if os.environ.get("MSD_LIB_PATH") else _HERE / "libmosdepth_b200.so"      # MSD_LIB_PATH: A/B of two builds in one GPU session

MODE_DEFAULT, MODE_FAST, MODE_FRAGMENT = 0, 1, 2

EXPORTS = [
    "msd_abi_version", "msd_create", "msd_destroy", "msd_last_error", "msd_open", "msd_open_mem", "msd_stage",
    "msd_set_spans", "msd_set_targets", "msd_set_filters", "msd_run", "msd_last_run_info", "msd_contig_records",
    "msd_contig_stats", "msd_n_windows", "msd_prefetch_windows", "msd_windows", "msd_regions", "msd_per_base_runs", "msd_per_base_bgzf", "msd_quantized_runs", "msd_quantized_bgzf",
    "msd_set_contig_range", "msd_set_contig_cover", "msd_contig_span_need", "msd_contig_spill", "msd_depth_copy", "msd_depth_add", "msd_contig_finalize",
    "msd_totals_device", "msd_totals_reset", "msd_comm_unique_id", "msd_comm_init", "msd_comm_destroy", "msd_exchange_spills", "msd_totals_allreduce", "msd_totals", "msd_mark", "msd_elapsed_ms", "msd_launch_count", "msd_debug_inflate", "msd_debug_depth", "msd_debug_cumsum", "msd_debug_crc32",
]


class DepthStat(C.Structure):
    _fields_ = [("cum_length", C.c_int64), ("cum_depth", C.c_uint64), ("min_depth", C.c_uint32), ("max_depth", C.c_uint32)]


class RunInfo(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_inflate", C.c_double), ("ms_frame", C.c_double),
                ("ms_records", C.c_double), ("ms_scan", C.c_double), ("ms_zero", C.c_double),
                ("bytes_compressed", C.c_uint64), ("bytes_inflated", C.c_uint64), ("n_blocks", C.c_uint64),
                ("n_records", C.c_uint64), ("n_events", C.c_uint64), ("n_depth_cells", C.c_uint64),
                ("n_launches", C.c_uint32), ("n_inflate_launches", C.c_uint32), ("ms_inflate_span", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MsdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmosdepth_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Load libmosdepth_b200.so (fails loudly when it has not been built: run `make` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} not built -- run `make` (or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    p, i32, i64, u64, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32
    lib.msd_abi_version.restype = i32
    lib.msd_create.argtypes = [i32, C.POINTER(p)]
    lib.msd_destroy.argtypes = [p]; lib.msd_destroy.restype = None
    lib.msd_last_error.argtypes = [p]; lib.msd_last_error.restype = C.c_char_p
    lib.msd_open.argtypes = [p, C.c_char_p, i32, p, u64]
    lib.msd_open_mem.argtypes = [p, p, u64, i32, p, u64]
    lib.msd_stage.argtypes = [p]
    lib.msd_set_spans.argtypes = [p, i32, p, p]
    lib.msd_set_targets.argtypes = [p, p]
    lib.msd_set_filters.argtypes = [p, i32, i64, i64, C.c_uint16, C.c_uint16, C.POINTER(C.c_char_p), i32, i32]
    lib.msd_run.argtypes = [p]
    lib.msd_last_run_info.argtypes = [p, C.POINTER(RunInfo)]
    lib.msd_contig_records.argtypes = [p, i32, C.POINTER(i64)]
    lib.msd_contig_stats.argtypes = [p, i32, C.POINTER(DepthStat), p, i64, C.POINTER(i64)]
    lib.msd_n_windows.argtypes = [u32, u32]; lib.msd_n_windows.restype = i64
    lib.msd_prefetch_windows.argtypes = [p, u32]
    lib.msd_set_contig_range.argtypes = [p, i32, u32, u32]
    lib.msd_set_contig_cover.argtypes = [p, i32, u32]
    lib.msd_contig_span_need.argtypes = [p, i32, C.POINTER(u32), C.POINTER(u32)]
    lib.msd_comm_unique_id.argtypes = [p]
    lib.msd_comm_init.argtypes = [p, i32, i32, p]
    lib.msd_comm_destroy.argtypes = [p]
    lib.msd_exchange_spills.argtypes = [p]
    lib.msd_totals_allreduce.argtypes = [p]
    lib.msd_totals.argtypes = [p, C.POINTER(DepthStat), C.POINTER(DepthStat), p, i64, C.POINTER(i64), p, i64, C.POINTER(i64)]
    lib.msd_contig_spill.argtypes = [p, i32, C.POINTER(u32), C.POINTER(u32)]
    lib.msd_depth_copy.argtypes = [p, i32, u32, u32, p, i32]
    lib.msd_depth_add.argtypes = [p, i32, u32, u32, p, i32]
    lib.msd_contig_finalize.argtypes = [p, i32]
    lib.msd_windows.argtypes = [p, i32, u32, i32, p, i64, C.POINTER(DepthStat), p]; lib.msd_windows.restype = i64
    lib.msd_regions.argtypes = [p, i32, i64, p, p, i32, p, C.POINTER(DepthStat), p, i64, C.POINTER(i64), p, i32, p]
    lib.msd_per_base_runs.argtypes = [p, i32, C.POINTER(p), C.POINTER(p), C.POINTER(p), C.POINTER(i64)]
    lib.msd_per_base_bgzf.argtypes = [p, i32, C.c_char_p, C.POINTER(p), C.POINTER(i64), C.POINTER(p), C.POINTER(i64), C.POINTER(p), C.POINTER(p), C.POINTER(p), C.POINTER(i64)]
    lib.msd_quantized_bgzf.argtypes = [p, i32, p, i32, C.c_char_p, C.POINTER(C.c_char_p), i32, C.POINTER(p), C.POINTER(i64), C.POINTER(p), C.POINTER(i64), C.POINTER(p), C.POINTER(p), C.POINTER(p), C.POINTER(i64)]
    lib.msd_quantized_runs.argtypes = [p, i32, p, i32, C.POINTER(p), C.POINTER(p), C.POINTER(p), C.POINTER(i64)]
    lib.msd_totals_device.argtypes = [p, C.POINTER(p), C.POINTER(p), C.POINTER(i64), C.POINTER(p)]
    lib.msd_totals_reset.argtypes = [p]
    lib.msd_mark.argtypes = [p, i32]
    lib.msd_elapsed_ms.argtypes = [p, i32, i32, C.POINTER(C.c_double)]
    lib.msd_launch_count.argtypes = [p]; lib.msd_launch_count.restype = u64
    lib.msd_debug_inflate.argtypes = [p, p, u64, p, u64]; lib.msd_debug_inflate.restype = i64
    lib.msd_debug_depth.argtypes = [p, i32, p, i64]
    lib.msd_debug_cumsum.argtypes = [p, p, i64]
    lib.msd_debug_crc32.argtypes = [p, p, u64, u32, u32, i32, p, C.POINTER(C.c_double)]; lib.msd_debug_crc32.restype = i64
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def comm_unique_id() -> bytes:
    """The 128-byte NCCL id rank 0 hands to every rank's Context.comm_init (msd_comm_unique_id)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.msd_comm_unique_id(buf)
    if rc < 0:
        raise MsdError(rc, lib.msd_last_error(None).decode())
    return buf.raw


class Context:
    """One msd_ctx on one GPU. Mirrors the call order of the reference's main loop (mosdepth.nim:589-840)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.msd_create(device, C.byref(h))
        if rc != 0:
            raise MsdError(rc, self.lib.msd_last_error(None).decode())
        self.h = h
        self._keep = []
        self.tlen = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.msd_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise MsdError(rc, self.lib.msd_last_error(self.h).decode())
        return rc

    # -- input ------------------------------------------------------------------------------------------------
    def open(self, path, header: "hostio.BamHeader | None" = None):
        header = header or hostio.read_bam_header(path)
        self.header = header
        self.tlen = np.asarray(header.lengths, dtype=np.uint32)
        self._check(self.lib.msd_open(self.h, str(path).encode(), len(self.tlen), _ptr(self.tlen), header.first_record_voff))
        return header

    def open_mem(self, buf, header: "hostio.BamHeader"):
        """buf: numpy uint8 array or an object with .data_ptr()/.numel() (a pinned torch tensor); kept alive here."""
        self.header = header
        self.tlen = np.asarray(header.lengths, dtype=np.uint32)
        if hasattr(buf, "data_ptr"):
            ptr, n = buf.data_ptr(), buf.numel()
        else:
            ptr, n = buf.ctypes.data, buf.size
        self._keep = [buf]
        self._check(self.lib.msd_open_mem(self.h, C.c_void_p(ptr), n, len(self.tlen), _ptr(self.tlen), header.first_record_voff))

    def stage(self):
        self._check(self.lib.msd_stage(self.h))

    def set_spans(self, begs, ends):
        b = np.asarray(begs, dtype=np.uint64); e = np.asarray(ends, dtype=np.uint64)
        self._check(self.lib.msd_set_spans(self.h, len(b), _ptr(b), _ptr(e)))

    def set_targets(self, want):
        w = None if want is None else np.asarray(want, dtype=np.uint8)
        self._check(self.lib.msd_set_targets(self.h, _ptr(w)))

    def set_filters(self, mapq=0, min_len=-1, max_len=-1, eflag=1796, iflag=0, read_groups=(), mode=MODE_DEFAULT):
        if max_len < 0:
            max_len = 2**63 - 1   # mosdepth.nim:929-930
        rgs = (C.c_char_p * max(1, len(read_groups)))(*[r.encode() for r in read_groups])
        self._check(self.lib.msd_set_filters(self.h, mapq, min_len, max_len, eflag, iflag, rgs, len(read_groups), mode))

    # -- run + queries ----------------------------------------------------------------------------------------------
    def run(self):
        self._check(self.lib.msd_run(self.h))
        return self.run_info()

    def run_info(self):
        info = RunInfo()
        self._check(self.lib.msd_last_run_info(self.h, C.byref(info)))
        return info

    def contig_records(self, tid):
        n = C.c_int64()
        rc = self.lib.msd_contig_records(self.h, tid, C.byref(n))
        if rc < -2:
            self._check(rc)
        return rc, n.value

    def contig_stats(self, tid):
        st = DepthStat(); n = C.c_int64()
        self._check(self.lib.msd_contig_stats(self.h, tid, C.byref(st), None, 0, C.byref(n)))
        dist = np.zeros(n.value, dtype=np.int64)
        self._check(self.lib.msd_contig_stats(self.h, tid, C.byref(st), _ptr(dist), dist.size, C.byref(n)))
        return st, dist

    # ---- intra-contig sharding (include/mosdepth_b200.h) ----
    def set_contig_range(self, tid, lo, hi):
        self._check(self.lib.msd_set_contig_range(self.h, tid, int(lo), int(hi)))

    def set_contig_cover(self, tid, covered_from):
        """Promise: the spans hold every record of contig tid with pos >= covered_from (default mode checks mates against it)."""
        self._check(self.lib.msd_set_contig_cover(self.h, tid, int(covered_from)))

    def contig_span_need(self, tid):
        a, b = C.c_uint32(), C.c_uint32()
        self._check(self.lib.msd_contig_span_need(self.h, tid, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- the collectives of the path (NCCL; one context per rank) ----
    def comm_unique_id(self):
        return comm_unique_id()

    def comm_init(self, rank, world, unique_id: bytes):
        self._check(self.lib.msd_comm_init(self.h, int(rank), int(world), C.c_char_p(unique_id)))

    def comm_destroy(self):
        self._check(self.lib.msd_comm_destroy(self.h))

    def exchange_spills(self):
        self._check(self.lib.msd_exchange_spills(self.h))

    def totals_allreduce(self):
        self._check(self.lib.msd_totals_allreduce(self.h))

    def totals(self):
        """(global DepthStat, region DepthStat, global dist, region dist) of the `total` rows -- msd_totals."""
        g, r = DepthStat(), DepthStat(); ng, nr = C.c_int64(), C.c_int64()
        self._check(self.lib.msd_totals(self.h, C.byref(g), C.byref(r), None, 0, C.byref(ng), None, 0, C.byref(nr)))
        gd = np.zeros(max(ng.value, 1), dtype=np.int64); rd = np.zeros(max(nr.value, 1), dtype=np.int64)
        self._check(self.lib.msd_totals(self.h, C.byref(g), C.byref(r), _ptr(gd), gd.size, C.byref(ng), _ptr(rd), rd.size, C.byref(nr)))
        return g, r, gd[:ng.value], rd[:nr.value]

    def contig_spill(self, tid):
        a, n = C.c_uint32(), C.c_uint32()
        self._check(self.lib.msd_contig_spill(self.h, tid, C.byref(a), C.byref(n)))
        return a.value, n.value

    def depth_copy(self, tid, pos, n, device_ptr=None):
        """Depth of [pos, pos + n) of a shard after run(): to a numpy array, or to device memory at device_ptr."""
        if device_ptr is not None:
            self._check(self.lib.msd_depth_copy(self.h, tid, int(pos), int(n), C.c_void_p(device_ptr), 1)); return None
        out = np.zeros(n, dtype=np.int32)
        self._check(self.lib.msd_depth_copy(self.h, tid, int(pos), int(n), _ptr(out), 0))
        return out

    def depth_add(self, tid, pos, cells=None, device_ptr=None, n=None):
        if device_ptr is not None:
            self._check(self.lib.msd_depth_add(self.h, tid, int(pos), int(n), C.c_void_p(device_ptr), 1)); return
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._check(self.lib.msd_depth_add(self.h, tid, int(pos), cells.size, _ptr(cells), 0))

    def contig_finalize(self, tid):
        self._check(self.lib.msd_contig_finalize(self.h, tid))

    def prefetch_windows(self, window):
        """Hint: msd_run() computes the windows of every contig in one launch; windows() then reads the host cache."""
        self._check(self.lib.msd_prefetch_windows(self.h, int(window)))

    def windows(self, tid, window, use_median=False):
        nw = self.lib.msd_n_windows(int(self.tlen[tid]), window)
        vals = np.zeros(max(nw, 1), dtype=np.float64); st = DepthStat(); dist = np.zeros(512, dtype=np.int64)
        got = self.lib.msd_windows(self.h, tid, window, int(use_median), _ptr(vals), vals.size, C.byref(st), _ptr(dist))
        self._check(got)
        return vals[:got], st, dist     # a shard of a contig (set_contig_range) returns only the windows that start in it

    def regions(self, tid, starts, stops, use_median=False, thresholds=()):
        s = np.ascontiguousarray(starts, dtype=np.uint32); e = np.ascontiguousarray(stops, dtype=np.uint32)
        n = len(s)
        vals = np.zeros(max(n, 1), dtype=np.float64); st = DepthStat(); dl = C.c_int64()
        dist = np.zeros(400001, dtype=np.int64)
        thr = np.ascontiguousarray(thresholds, dtype=np.int32)
        cnt = np.zeros((max(n, 1), max(len(thr), 1)), dtype=np.int64)
        self._check(self.lib.msd_regions(self.h, tid, n, _ptr(s), _ptr(e), int(use_median), _ptr(vals), C.byref(st), _ptr(dist), dist.size,
                                         C.byref(dl), _ptr(thr) if len(thr) else None, len(thr), _ptr(cnt) if len(thr) else None))
        return vals[:n], st, dist[:dl.value], cnt[:n, :len(thr)]

    def _runs(self, fn, tid, *extra, copy=True):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p(); n = C.c_int64()
        self._check(fn(self.h, tid, *extra, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        def arr(p):
            if n.value == 0:
                return np.zeros(0, dtype=np.int32)
            v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value,))
            return v.copy() if copy else v      # copy=False: a view of the library-owned pinned buffer, valid until the next call that returns runs
        return arr(a), arr(b), arr(c)

    def per_base_runs(self, tid, copy=True):
        return self._runs(self.lib.msd_per_base_runs, tid, copy=copy)

    def per_base_bgzf(self, tid, chrom):
        """(BGZF bytes without the EOF marker, members as a structured array, (start, stop, depth)) -- msd_per_base_bgzf."""
        return self._bgzf(lambda *out: self.lib.msd_per_base_bgzf(self.h, tid, chrom.encode() if isinstance(chrom, str) else chrom, *out))

    def quantized_bgzf(self, tid, quants, chrom, labels):
        """The same for quantized.bed.gz: labels[bin] is the 4th field -- msd_quantized_bgzf."""
        q = np.ascontiguousarray(quants, dtype=np.int64)
        lab = (C.c_char_p * max(1, len(labels)))(*[l.encode() if isinstance(l, str) else l for l in labels])
        return self._bgzf(lambda *out: self.lib.msd_quantized_bgzf(self.h, tid, _ptr(q), len(q), chrom.encode() if isinstance(chrom, str) else chrom, lab, len(labels), *out))

    def _bgzf(self, call):
        z, m, a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(); nz, nm, n = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(call(C.byref(z), C.byref(nz), C.byref(m), C.byref(nm), C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        def arr(p):
            if n.value == 0:
                return np.zeros(0, dtype=np.int32)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value,)).copy()
        data = C.string_at(z, nz.value) if nz.value else b""
        mdt = np.dtype([("offset", "<u8"), ("first_run", "<u8"), ("csize", "<u4"), ("isize", "<u4")])
        members = np.frombuffer(C.string_at(m, nm.value * mdt.itemsize), dtype=mdt).copy() if nm.value else np.zeros(0, dtype=mdt)
        return data, members, (arr(a), arr(b), arr(c))

    def quantized_runs(self, tid, quants, copy=True):
        q = np.ascontiguousarray(quants, dtype=np.int64)
        return self._runs(self.lib.msd_quantized_runs, tid, _ptr(q), len(q), copy=copy)

    def totals_device(self):
        g, r, s = C.c_void_p(), C.c_void_p(), C.c_void_p(); n = C.c_int64()
        self._check(self.lib.msd_totals_device(self.h, C.byref(g), C.byref(r), C.byref(n), C.byref(s)))
        return g.value, r.value, n.value, s.value

    def totals_reset(self):
        self._check(self.lib.msd_totals_reset(self.h))

    def mark(self, slot):
        self._check(self.lib.msd_mark(self.h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_double()
        self._check(self.lib.msd_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.msd_launch_count(self.h))

    # -- test hooks -------------------------------------------------------------------------------------------------
    def debug_inflate(self, data: bytes, out_size: int):
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(max(out_size, 1), dtype=np.uint8)
        n = self.lib.msd_debug_inflate(self.h, _ptr(src), src.size, _ptr(out), out_size)
        self._check(n)
        return out[:n].tobytes()

    def debug_crc32(self, data, block_bytes=65280, first_offset=0, reps=0):
        """K1c alone: CRC-32 of every `block_bytes` piece of `data` (bytes or uint8 array); returns (crcs, ms per launch)."""
        src = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
        nb = (src.size + block_bytes - 1) // block_bytes
        out = np.zeros(max(nb, 1), dtype=np.uint32); ms = C.c_double()
        got = self.lib.msd_debug_crc32(self.h, _ptr(src), src.size, block_bytes, first_offset, reps, _ptr(out), C.byref(ms))
        self._check(got)
        return out[:got], ms.value

    def debug_depth(self, tid):
        out = np.zeros(int(self.tlen[tid]) + 1, dtype=np.int32)
        self._check(self.lib.msd_debug_depth(self.h, tid, _ptr(out), out.size))
        return out

    def debug_cumsum(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        self._check(self.lib.msd_debug_cumsum(self.h, _ptr(a), a.size))
        return a
