"""Shared ctypes loaders and data helpers for the test-suite (test infrastructure only).

* ``oracle()``    -> oracle/liboracle.so   (CPU restatement; built by ``make -C oracle oracle``)
* ``ref_shim()``  -> oracle/_ref/libref_shim.so (the unmodified reference behind a C shim; only
                     present when it was built in the container that has /root/reference)
* ``product()``   -> fasim-longtarget_b200/libfasim_b200.so (the CUDA product, C ABI of include/fasim_b200.h)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
PKG_DIR = os.path.join(ROOT, "fasim-longtarget_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")

DEFAULT_PARAMS = dict(rule=0, cutLength=5000, strand=0, overlap=100, ntMin=20, ntMax=100000, minIdentity=60,
                      minStability=1, penaltyT=-1000, penaltyC=0, cDistance=15, cLength=50)
PARAM_ORDER = ["rule", "cutLength", "strand", "overlap", "ntMin", "ntMax", "minIdentity", "minStability", "penaltyT",
               "penaltyC", "cDistance", "cLength"]

# reference task order (Fasim-LongTarget.cpp:404-585): (para, strand, rule)
TASKS = [(1, s, r) for r in range(1, 7) for s in (0, 1)] + [(-1, s, r) for r in range(1, 19) for s in (1, 0)]


def params_array(**kw):
    p = dict(DEFAULT_PARAMS)
    p.update(kw)
    return (C.c_int * 12)(*[int(p[k]) for k in PARAM_ORDER])


_cache = {}


def _load(path):
    if path not in _cache:
        _cache[path] = C.CDLL(path)
    return _cache[path]


def oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"])
    return _load(path)


def have_ref_shim():
    return os.path.exists(os.path.join(REF_DIR, "libref_shim.so"))


def ref_shim():
    return _load(os.path.join(REF_DIR, "libref_shim.so"))


def ref_binary():
    return os.path.join(REF_DIR, "fasim")


class Side:
    """Uniform python face over either the oracle (prefix 'orc_') or the reference shim (prefix 'ref_')."""

    def __init__(self, lib, prefix):
        self.lib, self.p = lib, prefix

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def task_strings(self, seg, para, strand, rule):
        a = C.create_string_buffer(len(seg) + 1)
        b = C.create_string_buffer(len(seg) + 1)
        self._f("task_strings")(seg.encode(), para, strand, rule, a, b)
        return a.value.decode(), b.value.decode()

    def calc_score_once(self, rna, seq2):
        return self._f("calc_score_once")(rna.encode(), seq2.encode())

    def colmax(self, rna, seq2):
        n = len(seq2)
        out = (C.c_int * max(n, 1))()
        self._f("colmax")(rna.encode(), seq2.encode(), n, out)
        return np.array(out[:n], dtype=np.int32)

    def prealign(self, rna, seq2, thr):
        cap = len(seq2) + 1
        s = (C.c_int * cap)()
        p = (C.c_int * cap)()
        k = self._f("prealign")(rna.encode(), seq2.encode(), len(seq2), thr, s, p, cap)
        return list(zip(s[:k], p[:k]))

    def align(self, rna, win):
        out6 = (C.c_int * 6)()
        cig = (C.c_uint * 4096)()
        self._f("align")(rna.encode(), win.encode(), len(win), out6, cig, 4096)
        return tuple(out6[:5]), list(cig[:out6[5]])

    def task(self, rna, seg, dna_start, para, strand, rule, **kw):
        cap = 1 << 22
        buf = C.create_string_buffer(cap)
        ms = C.c_int(0)
        f = self._f("task")
        f.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                      C.c_char_p, C.c_long]
        k = f(rna.encode(), seg.encode(), dna_start, para, strand, rule, params_array(**kw), C.byref(ms), buf, cap)
        assert k >= 0
        return ms.value, buf.value.decode()

    def sim_task(self, rna, seg, dna_start, para, strand, rule, **kw):
        """one task in -F mode: calc_score_once + SIM (Fasim-LongTarget.cpp:419-426, sim.h:410)"""
        cap = 1 << 22
        buf = C.create_string_buffer(cap)
        ms = C.c_int(0)
        f = self._f("sim_task")
        f.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                      C.c_char_p, C.c_long]
        k = f(rna.encode(), seg.encode(), dna_start, para, strand, rule, params_array(**kw), C.byref(ms), buf, cap)
        assert k >= 0
        return ms.value, buf.value.decode()

    def sim_longtarget(self, rna, dna, **kw):
        cap = 1 << 26
        buf = C.create_string_buffer(cap)
        f = self._f("sim_longtarget")
        f.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_long]
        k = f(rna.encode(), dna.encode(), params_array(**kw), buf, cap)
        assert k >= 0
        return buf.value.decode()

    def longtarget(self, rna, dna, **kw):
        cap = 1 << 26
        buf = C.create_string_buffer(cap)
        f = self._f("longtarget")
        f.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_long]
        k = f(rna.encode(), dna.encode(), params_array(**kw), buf, cap)
        assert k >= 0
        return buf.value.decode()

    def run_lowercase(self, rna, records, **kw):
        """oracle only: the older variant's whole run (fasim-LongTarget.cpp) -> (text of -fastSim-TFOsorted, out-of-bounds columns);
        records = [(dna, chr, start)]"""
        n = len(records)
        cap = 1 << 26
        buf = C.create_string_buffer(cap)
        oob = C.c_long(0)
        f = self._f("run_lowercase")
        f.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_long), C.POINTER(C.c_int), C.c_char_p,
                      C.c_long, C.POINTER(C.c_long)]
        k = f(rna.encode(), n, (C.c_char_p * n)(*[d.encode() for d, _, _ in records]), (C.c_char_p * n)(*[c.encode() for _, c, _ in records]),
              (C.c_long * n)(*[int(s) for _, _, s in records]), params_array(**kw), buf, cap, C.byref(oob))
        assert k >= 0
        return buf.value.decode(), oob.value

    def cluster(self, stari, endi, nt, dd, length):
        n = len(stari)
        arr = lambda v: (C.c_int * n)(*[int(x) for x in v])
        mid, cen, mot = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        self._f("cluster")(n, arr(stari), arr(endi), arr(nt), dd, length, mid, cen, mot)
        return list(mid), list(cen), list(mot)


def oracle_side():
    return Side(oracle(), "orc_")


def ref_side():
    return Side(ref_shim(), "ref_")


def read_fasta(path):
    """-> list of (header_without_gt, sequence)"""
    recs, name, parts = [], None, []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(parts)))
                name, parts = line[1:], []
            else:
                parts.append(line)
    if name is not None:
        recs.append((name, "".join(parts)))
    return recs


def splitmix_bases(seed, n):
    """SURVEY.md §8(d) generator: SplitMix64, base = 'ACGT'[z >> 62] — vectorised."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        s = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[(z >> np.uint64(62)).astype(np.int64)].tobytes().decode()
