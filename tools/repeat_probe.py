#!/usr/bin/env python
"""Throughput on repeat-rich DNA (not the headline workload): microsatellites, purine tracts and a dispersed repeat family
planted into SplitMix64 background at a given density, against an lncRNA with (CT)/(GA)/(GT)-rich tracts.  Prints GCUPS,
stage times and the number of tasks that needed the literal (Q4) emulation.  Usage: repeat_probe.py [Mbp] [plants per 100 kb]"""
import ctypes as C
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200"))
sys.path.insert(0, ROOT)
import fasim_b200 as fb
from bench import splitmix_bases

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
density = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n = int(mbp * 1e6)
rng = random.Random(7)


def mutate(s, rate):
    return "".join(rng.choice("ACGT") if rng.random() < rate else c for c in s)


rna = bytearray(splitmix_bases(2001, 3000).tobytes())
for at, unit in ((500, "CT"), (1500, "GA"), (2500, "GT")):
    rna[at:at + 200] = mutate(unit * 100, 0.08).encode()
rna = rna.decode()
dna = bytearray(splitmix_bases(1001, n).tobytes())
alu = splitmix_bases(555, 150).tobytes().decode() + mutate("GA" * 40, 0.05) + splitmix_bases(556, 70).tobytes().decode()
for _ in range(int(n / 1e5 * density)):
    at = rng.randrange(0, n - 400)
    if rng.random() < 0.25:
        s = mutate(alu, 0.06)
    else:
        unit = rng.choice(["GA", "CT", "GAA", "CCT", "A", "GGA", "TC", "AG"])
        L = rng.randrange(30, 260)
        s = mutate((unit * L)[:L], rng.choice([0.0, 0.05, 0.12]))
    dna[at:at + len(s)] = s.encode()
dna = bytes(dna)
eng = fb.Engine(0)
eng.set_query("lnc", rna)
for it in range(2):
    res = C.POINTER(fb.Result)()
    t0 = time.perf_counter()
    rc = fb.lib().ltg_scan_record(eng._h, dna, len(dna), b"chr1", 1, C.byref(res))
    dt = time.perf_counter() - t0
    assert rc == 0, fb.lib().ltg_last_error()
    r = res.contents
    out = {"mbp": mbp, "plants_per_100kb": density, "seconds": dt, "gcups": r.scan_cells / dt / 1e9, "rows": r.n_triplex, "peaks": r.n_peaks,
           "scan_ms": r.gpu_ms_scan_kernel, "window_ms": r.gpu_ms_window, "literal_tasks": r.n_literal_tasks, "q4_probed_pairs": r.n_q4_probed, "literal_windows": r.n_literal_windows,
           "window_cells_over_scan_cells": r.window_cells / max(r.scan_cells, 1)}
    fb.lib().ltg_result_free(res)
print(json.dumps(out))
print(json.dumps(eng.debug_stats()), file=sys.stderr)
