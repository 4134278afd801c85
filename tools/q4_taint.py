#!/usr/bin/env python
"""Research prototype (CPU, numpy): certify a task as unaffected by the reference's Q4 quirk WITHOUT emulating the 8-bit kernel.
The exact DP carries one extra bit per value (value * 2 + 1 = "also reachable without anything the quirk can lose"; a maximum
prefers the untainted alternative on ties).  Contributions of a cross-stripe F chain are tainted from the row after the chain
passes through [132, 143] (the only values at which the signed lazy-F test can end the loop early).  A task is certified when no
recorded column maximum is tainted and no tainted chain >= 132 exists.  See DESIGN.md 7.
Usage: q4_taint.py synth [cases] | q4_taint.py data <lncRNA> [regions]"""
import gzip, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _harness import ref_side
from test_q4_theory_cpu import exact_colmax_and_carried_f, make_case
MATCH, MISMATCH, OPEN, EXT = 10, -8, 32, 8            # doubled scores: bit 0 is the "untainted" flag


def certify(rna, dna):
    m, n = len(rna), len(dna)
    L = (m + 15) // 16; m16 = 16 * L
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 0}
    d = np.array([code.get(c, 4) for c in dna])
    idx = np.arange(n)
    one = np.ones(n, dtype=np.int64)
    H = one.copy(); Hmain_prev = one.copy()
    fin = one.copy(); fcar = one.copy()
    cut = np.zeros(n, dtype=bool)            # the chain of this column has passed through [132, 143]
    colmax = one.copy()
    giveup = False
    for i in range(m16):
        if i > 0:
            fend = np.maximum(np.maximum(fin - EXT, Hmain_prev - OPEN), 1)
            if i % L == 0:
                old = np.maximum(fcar - EXT, 1)
                newer = (fend >> 1) >= (old >> 1)
                fcar = np.where(newer, fend, old)
                cut = np.where(newer, False, cut_next)
                fin = one.copy()
                if np.any(((fcar & 1) == 0) & ((fcar >> 1) >= 132)):
                    giveup = True            # a chain that may itself be lower in the reference: its failing rows are unknown
            else:
                fin = fend
                fcar = np.maximum(fcar - EXT, 1)
                cut = cut_next
        v = fcar >> 1
        cut_next = cut | ((v >= 132) & (v <= 143))
        if i < m:
            r = code.get(rna[i], 4)
            s = np.where((d == r) & (d < 4), MATCH, MISMATCH) if r < 4 else np.full(n, MISMATCH)
        else:
            s = np.zeros(n, dtype=np.int64)
        diag = np.concatenate(([1], H[:-1]))
        t0 = np.maximum(diag + s, 1)
        pm = np.maximum.accumulate(t0 + EXT * idx)
        E = np.maximum(np.concatenate(([1], pm[:-1] - OPEN - EXT * (idx[1:] - 1))), 1)
        T = np.maximum(t0, E)
        Hmain = np.maximum(T, fin)
        contrib = np.where(cut, fcar & ~1, fcar)
        H = np.maximum(Hmain, contrib)
        Hmain_prev = Hmain
        colmax = np.maximum(colmax, H)
    val = colmax >> 1
    over = np.nonzero(val >= 251)[0]
    jstar = int(over[0]) if len(over) else n
    je = min(jstar + 1, n)
    clean = bool(np.all((colmax[:je] & 1) == 1))
    return clean and not giveup


def run(cases):
    S = ref_side()
    st = dict(cases=0, flagged=0, certified=0, really_different=0, certified_but_different=0)
    for rna, dna in cases:
        exact, fmax = exact_colmax_and_carried_f(rna, dna)
        st["cases"] += 1
        if fmax < 132:
            continue
        st["flagged"] += 1
        differs = not np.array_equal(exact, S.colmax(rna, dna))
        ok = certify(rna, dna)
        st["certified"] += int(ok); st["really_different"] += int(differs); st["certified_but_different"] += int(ok and differs)
    return st


if sys.argv[1] == "synth":
    rng = random.Random(5)
    print(run(make_case(rng) for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1000)))
else:
    name = sys.argv[2]; n_regions = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    D = os.path.join(ROOT, "tests", "golden", "data")
    rna = "".join(l.strip() for l in open(os.path.join(D, name + ".fa")).read().splitlines()[1:])
    recs, cur = [], None
    for line in gzip.open(os.path.join(D, "MEG3-DNAseq.fa.gz"), "rt"):
        if line.startswith(">"):
            cur = []; recs.append(cur)
        elif cur is not None:
            cur.append(line.strip())
    recs = ["".join(r) for r in recs]
    rng = random.Random(3)
    S = ref_side()
    TASKS = [(1, s, r) for r in range(1, 7) for s in (0, 1)] + [(-1, s, r) for r in range(1, 19) for s in (0, 1)]
    print(run((rna, S.task_strings(recs[reg][:5000], p, s, r)[0]) for reg in rng.sample(range(len(recs)), n_regions) for p, s, r in TASKS))
