"""ctypes face of libfasim_b200.so — host-side mirror of the reference's interface for the hot path.

The names follow the reference (Fasim-LongTarget.cpp / fastsim.h / ssw_cpp.h): ``calc_score_once``, ``preAlign``,
``Align``, ``LongTarget``, ``cluster_triplex``, ``printResult``; every call goes through the C ABI declared in
include/fasim_b200.h and therefore through the CUDA kernels.  There is no CPU fallback: importing works anywhere,
creating an :class:`Engine` needs a B200.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FASIM_B200_LIB") or os.path.join(HERE, "libfasim_b200.so")     # FASIM_B200_LIB: a tuning variant
CLI_PATH = os.path.join(HERE, "fasim")

PARAM_FIELDS = ["rule", "cut_length", "strand", "overlap", "nt_min", "nt_max", "min_identity", "min_stability", "penalty_t",
                "penalty_c", "c_distance", "c_length"]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in PARAM_FIELDS]


class Triplex(C.Structure):
    _fields_ = [("stari", C.c_int32), ("endi", C.c_int32), ("starj", C.c_int32), ("endj", C.c_int32), ("reverse", C.c_int32),
                ("strand", C.c_int32), ("rule", C.c_int32), ("nt", C.c_int32), ("score", C.c_float), ("identity", C.c_float),
                ("tri_score", C.c_float), ("middle", C.c_int32), ("center", C.c_int32), ("motif", C.c_int32),
                ("genomestart", C.c_int64), ("genomeend", C.c_int64), ("tfo_off", C.c_int64), ("tts_off", C.c_int64),
                ("chr_off", C.c_int64), ("record", C.c_int32), ("pad_", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("n_triplex", C.c_int64), ("triplex", C.POINTER(Triplex)), ("text_bytes", C.c_int64), ("text", C.POINTER(C.c_char)),
                ("n_segments", C.c_int64), ("n_tasks", C.c_int64), ("scan_cells", C.c_int64), ("dna_bases", C.c_int64),
                ("n_peaks", C.c_int64), ("window_cells", C.c_int64), ("n_literal_tasks", C.c_int64),
                ("n_literal_windows", C.c_int64), ("gpu_ms_scan", C.c_double), ("gpu_ms_window", C.c_double),
                ("gpu_launches", C.c_int64), ("gpu_ms_scan_kernel", C.c_double), ("n_scan_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("n_q4_probed", C.c_int64)]


class TaskProbe(C.Structure):
    _fields_ = [("para", C.c_int32), ("strand", C.c_int32), ("rule", C.c_int32), ("max_score", C.c_int32), ("threshold", C.c_int32),
                ("n_peaks", C.c_int32), ("literal", C.c_int32), ("pad_", C.c_int32)]


EXPORTS = ["ltg_create", "ltg_destroy", "ltg_last_error", "ltg_default_params", "ltg_set_params", "ltg_set_sim_mode", "ltg_set_compat", "ltg_set_query",
           "ltg_scan_record", "ltg_scan_records", "ltg_scan_records_at", "ltg_scan_device", "ltg_scan_shard", "ltg_scan_packed", "ltg_result_append", "ltg_result_new", "ltg_result_free", "ltg_cluster",
           "ltg_write_tfosorted", "ltg_write_tfoclass", "ltg_main", "ltg_probe_segment", "ltg_probe_align", "ltg_stream", "ltg_debug_stats",
           "ltg_device_count"]

_lib = None


def build(force=False):
    """Compile libfasim_b200.so + the fasim CLI for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH):
        subprocess.check_call(["sh", os.path.join(HERE, "build.sh")])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libfasim_b200.so is missing — run fasim-longtarget_b200/build.sh (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ltg_last_error.restype = C.c_char_p
        L.ltg_stream.restype = C.c_void_p
        L.ltg_stream.argtypes = [C.c_void_p]
        L.ltg_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.ltg_destroy.argtypes = [C.c_void_p]
        L.ltg_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.ltg_set_sim_mode.argtypes = [C.c_void_p, C.c_int]
        L.ltg_set_compat.argtypes = [C.c_void_p, C.c_int]
        L.ltg_set_query.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int64]
        L.ltg_scan_record.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.POINTER(Result))]
        L.ltg_scan_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.POINTER(Result))]
        L.ltg_scan_records.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_char_p),
                                       C.POINTER(C.c_int64), C.POINTER(C.POINTER(Result))]
        L.ltg_scan_records_at.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_char_p),
                                          C.POINTER(C.c_int64), C.POINTER(C.POINTER(Result))]
        L.ltg_scan_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                     C.POINTER(C.POINTER(Result))]
        L.ltg_scan_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int32,
                                      C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.POINTER(Result))]
        L.ltg_result_free.argtypes = [C.POINTER(Result)]
        L.ltg_result_new.argtypes = [C.POINTER(C.POINTER(Result))]
        L.ltg_result_append.argtypes = [C.POINTER(Result), C.POINTER(Result)]
        L.ltg_cluster.argtypes = [C.POINTER(Result), C.POINTER(Params)]
        L.ltg_write_tfosorted.argtypes = [C.POINTER(Result), C.c_char_p]
        L.ltg_write_tfoclass.argtypes = [C.POINTER(Result), C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_char_p]
        L.ltg_probe_segment.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(TaskProbe), C.c_int32, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
        L.ltg_probe_align.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_uint32), C.c_int32]
        L.ltg_debug_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        _lib = L
    return _lib


class FasimError(RuntimeError):
    pass


def default_params(**kw):
    p = Params()
    lib().ltg_default_params(C.byref(p))
    for k, v in kw.items():
        if k not in PARAM_FIELDS:
            raise KeyError(k)
        setattr(p, k, int(v))
    return p


def _check(rc):
    if rc != 0:
        raise FasimError("libfasim_b200 error %d: %s" % (rc, lib().ltg_last_error().decode()))


def result_rows(res):
    """ltg_result -> list of dicts (floats as python floats of the float32 values)."""
    r = res.contents
    text = C.string_at(r.text, r.text_bytes) if r.text_bytes else b""

    def s(off):
        end = text.index(b"\0", off)
        return text[off:end].decode()

    out = []
    for i in range(r.n_triplex):
        t = r.triplex[i]
        d = {f[0]: getattr(t, f[0]) for f in Triplex._fields_ if not f[0].endswith("_off") and f[0] != "pad_"}
        d["tfo"], d["tts"], d["chr"] = s(t.tfo_off), s(t.tts_off), s(t.chr_off)
        out.append(d)
    return out


class Engine:
    """One GPU context (reference: one `fasim` process)."""

    def __init__(self, device=0, **params):
        self._h = C.c_void_p()
        _check(lib().ltg_create(device, C.byref(self._h)))
        self.params = default_params(**params)
        _check(lib().ltg_set_params(self._h, C.byref(self.params)))
        self.rna = None

    def close(self):
        if self._h:
            lib().ltg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **params):
        self.params = default_params(**params)
        _check(lib().ltg_set_params(self._h, C.byref(self.params)))

    def set_sim_mode(self, on=True):
        """-F of the reference (Fasim-LongTarget.cpp:360): SIM() instead of fastSIM() per task."""
        _check(lib().ltg_set_sim_mode(self._h, 1 if on else 0))

    def set_compat(self, lowercase=True):
        """the older variant's per-task pipeline (fasim-LongTarget.cpp / fastSim.h)"""
        _check(lib().ltg_set_compat(self._h, 1 if lowercase else 0))

    def set_query(self, name, rna):
        self.rna = rna
        b = rna.encode()
        _check(lib().ltg_set_query(self._h, name.encode(), b, len(b)))

    def debug_stats(self, reset=False):
        out = (C.c_int64 * 30)()
        lib().ltg_debug_stats(self._h, out, 1 if reset else 0)
        return {"windows": list(out[:10]), "cells": list(out[10:20]), "traceback_handed_over": list(out[20:24]),
                "filters(live,nt_bound_fail,nt_ok,identity_ok,stability_ok,all_ok)": list(out[24:30])}

    @property
    def stream(self):
        return lib().ltg_stream(self._h)

    # ---- reference-shaped calls ---------------------------------------------------------------
    def probe_segment(self, seg, tasks, want_colmax=True, peak_cap=2048):
        """tasks: list of (para, strand, rule).  -> list of dicts with max_score (calc_score_once, stats.h:879), threshold,
        colmax (ssw_pre_align, sswNew.cpp:1309), peaks (Aligner::preAlign, ssw_cpp.cpp:388) and the literal flag."""
        import numpy as np
        n, k = len(seg), len(tasks)
        arr = (TaskProbe * k)()
        for i, (pa, st, ru) in enumerate(tasks):
            arr[i].para, arr[i].strand, arr[i].rule = pa, st, ru
        cm = np.zeros((k, n), dtype=np.int32) if want_colmax else None
        ps = np.zeros((k, peak_cap), dtype=np.int32)
        pp = np.zeros((k, peak_cap), dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        _check(lib().ltg_probe_segment(self._h, seg.encode(), n, arr, k, cm.ctypes.data_as(ip) if want_colmax else None,
                                       ps.ctypes.data_as(ip), pp.ctypes.data_as(ip), peak_cap))
        out = []
        for i in range(k):
            npk = min(arr[i].n_peaks, peak_cap)
            out.append(dict(max_score=arr[i].max_score, threshold=arr[i].threshold, n_peaks=arr[i].n_peaks, literal=arr[i].literal,
                            colmax=cm[i] if want_colmax else None, peaks=[(int(ps[i, j]), int(pp[i, j])) for j in range(npk)]))
        return out

    def calc_score_once(self, seg, para, strand, rule):
        return self.probe_segment(seg, [(para, strand, rule)], want_colmax=False)[0]["max_score"]

    def preAlign(self, seg, para, strand, rule):
        return self.probe_segment(seg, [(para, strand, rule)], want_colmax=False)[0]["peaks"]

    def Align(self, windows, cigar_cap=256):
        """windows: translated-DNA strings.  -> list of ((sw_score, ref_begin, ref_end, query_begin, query_end), cigar)."""
        n = len(windows)
        enc = [w.encode() for w in windows]
        ptrs = (C.c_char_p * n)(*enc)
        lens = (C.c_int32 * n)(*[len(w) for w in enc])
        out6 = (C.c_int32 * (6 * n))()
        cig = (C.c_uint32 * (cigar_cap * n))()
        _check(lib().ltg_probe_align(self._h, ptrs, lens, n, out6, cig, cigar_cap))
        res = []
        for i in range(n):
            o = out6[6 * i:6 * i + 6]
            res.append((tuple(o[:5]), list(cig[i * cigar_cap:i * cigar_cap + min(o[5], cigar_cap)])))
        return res

    def scan_record(self, dna, chr_tag="", record_start=0):
        """LongTarget() for one record (Fasim-LongTarget.cpp:379) — returns the raw ltg_result pointer (free with .free)."""
        res = C.POINTER(Result)()
        b = dna.encode() if isinstance(dna, str) else dna
        _check(lib().ltg_scan_record(self._h, b, len(b), chr_tag.encode(), record_start, C.byref(res)))
        return res

    def scan_device(self, dev_ptr, length, chr_tag="", record_start=0):
        res = C.POINTER(Result)()
        _check(lib().ltg_scan_device(self._h, C.c_void_p(dev_ptr), length, chr_tag.encode(), record_start, C.byref(res)))
        return res

    def scan_records(self, records):
        """records: list of (dna, chr_tag, record_start) — all records of a multi-record FASTA in one call."""
        n = len(records)
        enc = [(d.encode() if isinstance(d, str) else d) for d, _, _ in records]
        tags = [t.encode() for _, t, _ in records]
        res = C.POINTER(Result)()
        _check(lib().ltg_scan_records(self._h, n, (C.c_char_p * n)(*enc), (C.c_int64 * n)(*[len(b) for b in enc]), (C.c_char_p * n)(*tags),
                                      (C.c_int64 * n)(*[int(s) for _, _, s in records]), C.byref(res)))
        return res

    def scan_shard(self, dna, record_len, first_segment, n_segments, chr_tag="", record_start=0, device_ptr=None):
        """Segments [first_segment, first_segment + n_segments) of a record of record_len bases; `dna` holds the bytes from
        the first segment's start (or pass device_ptr + length as `dna`)."""
        res = C.POINTER(Result)()
        if device_ptr is not None:
            _check(lib().ltg_scan_shard(self._h, C.c_void_p(device_ptr), 1, int(dna), chr_tag.encode(), record_start, record_len,
                                        first_segment, n_segments, C.byref(res)))
        else:
            b = dna.encode() if isinstance(dna, str) else dna
            _check(lib().ltg_scan_shard(self._h, C.cast(C.c_char_p(b), C.c_void_p), 0, len(b), chr_tag.encode(), record_start, record_len,
                                        first_segment, n_segments, C.byref(res)))
        return res

    def scan_packed(self, packed, first_base, n_bases, n_blocks=(), chr_tag="", record_start=0, record_len=-1, first_segment=0, n_segments=-1,
                    device_ptr=None):
        """2-bit packed DNA (UCSC coding) from host bytes or a device pointer; n_blocks = [(start, size)] relative to first_base."""
        res = C.POINTER(Result)()
        nb = len(n_blocks)
        ns = (C.c_uint32 * max(nb, 1))(*[int(a) for a, _ in n_blocks])
        nz = (C.c_uint32 * max(nb, 1))(*[int(b) for _, b in n_blocks])
        if device_ptr is not None:
            ptr, on_dev = C.c_void_p(device_ptr), 1
        else:
            self._packed_keep = bytes(packed)
            ptr, on_dev = C.cast(C.c_char_p(self._packed_keep), C.c_void_p), 0
        _check(lib().ltg_scan_packed(self._h, ptr, on_dev, first_base, n_bases, ns, nz, nb, chr_tag.encode(), record_start,
                                     n_bases if record_len < 0 else record_len, first_segment, n_segments, C.byref(res)))
        return res

    def LongTarget(self, dna, chr_tag="", record_start=0):
        res = self.scan_record(dna, chr_tag, record_start)
        rows = result_rows(res)
        lib().ltg_result_free(res)
        return rows

    def cluster_triplex(self, res):
        _check(lib().ltg_cluster(res, C.byref(self.params)))

    def printResult(self, res, path):
        _check(lib().ltg_write_tfosorted(res, path.encode()))

    @staticmethod
    def free(res):
        lib().ltg_result_free(res)


def shard_segments(n_bases, world, rank, cut=5000, overlap=100):
    """Contiguous shard of a record's segments for `rank` of `world` (cutSequence geometry, fastsim.h:71-90):
    -> (first_segment, n_segments, first_byte, n_bytes).  The byte range covers the shard's segments completely."""
    stride = cut - overlap
    n_seg = (n_bases + stride - 1) // stride if n_bases > 0 else 0
    lo, hi = (n_seg * rank) // world, (n_seg * (rank + 1)) // world
    if hi <= lo:
        return lo, 0, min(lo * stride, n_bases), 0
    first_byte = lo * stride
    last_byte = min(n_bases, (hi - 1) * stride + cut)
    return lo, hi - lo, first_byte, last_byte - first_byte


def merge_shard_rows(parts):
    """Concatenate per-rank row lists in rank (= segment) order: the order ltg_scan_record produces for the whole record."""
    out = []
    for rows in parts:
        out.extend(rows)
    return out


def run_cli(args, cwd=None):
    """Run the drop-in `fasim` binary (same flags as the reference)."""
    return subprocess.run([CLI_PATH] + list(args), cwd=cwd, capture_output=True, text=True)
