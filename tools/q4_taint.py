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
from test_q4_theory_cpu import certify


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
