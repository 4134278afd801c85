"""CPU: the -F mode (SIM, sim.h:410-1143).  Three layers, each pinned to the one before:
  reference shim (unmodified sim.h behind oracle/ref_shim.cpp)  ==  oracle restatement (oracle/oracle.cpp, namespace sim)
  ==  the product's sequential core (csrc/sim_core.cuh + host/sim_host.hpp, compiled for the host by tests/sim_core_host.cpp).
The device kernel (csrc/sim.cuh) shares that core; its wavefront first pass is checked on the GPU (tests/test_gpu_parity.py)."""
import ctypes as C
import os
import random
import subprocess

import pytest

from _harness import ROOT, TASKS, have_ref_shim, oracle_side, params_array, ref_side, splitmix_bases

O = oracle_side()


@pytest.fixture(scope="module")
def core():
    out = os.path.join(ROOT, "tests", "_libsimcore_host.so")
    src = os.path.join(ROOT, "tests", "sim_core_host.cpp")
    deps = [src] + [os.path.join(ROOT, "fasim-longtarget_b200", p) for p in ("csrc/sim_core.cuh", "host/sim_host.hpp", "host/rules_table.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I/usr/local/cuda/include", "-o", out, src])
    lib = C.CDLL(out)
    lib.simcore_task.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_long]

    def run(rna, seg, dna_start, para, strand, rule, min_score, **kw):
        cap = 1 << 22
        buf = C.create_string_buffer(cap)
        k = lib.simcore_task(rna.encode(), seg.encode(), dna_start, para, strand, rule, min_score, params_array(**kw), buf, cap)
        assert k >= 0, k
        return buf.value.decode()
    return run


def planted(trial, m, n, rnd):
    rna = splitmix_bases(100 + trial, m)
    dna = list(splitmix_bases(200 + trial, n))
    for _ in range(3):
        L = rnd.randrange(25, min(70, m - 1))
        at, a = rnd.randrange(0, n - L), rnd.randrange(0, m - L)
        dna[at:at + L] = rna[a:a + L].translate(str.maketrans("TG", "AT"))
    return rna, "".join(dna)


def noisy(rnd, unit, n, rate):
    s = list((unit * (n // len(unit) + 1))[:n])
    for i in range(n):
        if rnd.random() < rate:
            s[i] = rnd.choice("ACGT")
    return "".join(s)


def rows_wo_chr(txt):
    return [l.split("\t")[:19] for l in txt.splitlines()]


@pytest.mark.skipif(not have_ref_shim(), reason="reference shim not built (no /root/reference)")
def test_oracle_sim_equals_reference():
    R = ref_side()
    rnd = random.Random(5)
    rows = 0
    for trial, (m, n) in enumerate([(120, 400), (300, 700), (211, 1000)]):
        rna, dna = planted(trial, m, n, rnd)
        for (pa, st, ru) in TASKS[::5] + TASKS[1::7]:
            a = O.sim_task(rna, dna, 7, pa, st, ru, cLength=20)
            b = R.sim_task(rna, dna, 7, pa, st, ru, cLength=20)
            assert a == b, (trial, pa, st, ru)
            rows += len(a[1].splitlines())
    assert rows > 40
    # record level (cutSequence + all 48 tasks + final filter) on a repeat-rich record: many alignments per task, regions
    # recomputed next to each other, the 50-node list overflowing
    rna = splitmix_bases(2001, 60) + noisy(rnd, "CT", 70, 0.08) + splitmix_bases(2002, 40) + noisy(rnd, "GA", 50, 0.05)
    dna = splitmix_bases(1001, 150) + noisy(rnd, "GA", 260, 0.1) + splitmix_bases(1002, 100) + noisy(rnd, "TC", 150, 0.06)
    kw = dict(cLength=15, ntMin=10, cutLength=400, overlap=60)
    a, b = O.sim_longtarget(rna, dna, **kw), R.sim_longtarget(rna, dna, **kw)
    assert a == b and len(a.splitlines()) > 100


def test_product_core_equals_oracle(core):
    rnd = random.Random(17)
    rows = 0
    for trial, (m, n) in enumerate([(1, 1), (2, 40), (37, 5), (150, 600), (333, 450), (90, 1500)]):
        if m > 30 and n > 100:
            rna, dna = planted(trial, m, n, rnd)
        else:
            rna, dna = splitmix_bases(100 + trial, m), splitmix_bases(200 + trial, n)
        for (pa, st, ru) in TASKS:
            ms, txt = O.sim_task(rna, dna, 11, pa, st, ru, cLength=20)
            got = core(rna, dna, 11, pa, st, ru, ms, cLength=20)
            assert rows_wo_chr(got) == rows_wo_chr(txt), (trial, pa, st, ru)
            rows += len(txt.splitlines())
    assert rows > 150
    # repeat-rich: the list overflows, alignments are forbidden cell by cell, rectangles grow
    rna = splitmix_bases(2001, 50) + noisy(rnd, "CT", 80, 0.08) + splitmix_bases(2002, 40) + noisy(rnd, "GA", 60, 0.05)
    dna = splitmix_bases(1001, 120) + noisy(rnd, "GA", 300, 0.1) + splitmix_bases(1002, 80) + noisy(rnd, "TC", 200, 0.06)
    rows = 0
    for (pa, st, ru) in TASKS:
        for kw in (dict(cLength=20), dict(cLength=15, ntMin=10, ntMax=60, penaltyT=-3, penaltyC=2)):
            ms, txt = O.sim_task(rna, dna, 0, pa, st, ru, **kw)
            got = core(rna, dna, 0, pa, st, ru, ms, **kw)
            assert rows_wo_chr(got) == rows_wo_chr(txt), (pa, st, ru, kw)
            rows += len(txt.splitlines())
    assert rows > 1000
