#!/usr/bin/env python
"""How often does the reference's Q4 quirk really change a task's column maxima on the example data?  CPU only (needs the
reference shim, i.e. this container): samples (region, rule, strand, orientation) tasks of <lncRNA> x the MEG3 example regions,
computes the exact column maxima + the carried-F criterion (numpy) and the reference's column maxima (shim).
Usage: q4_rate.py H19|MALAT1|MEG3-ENST00000451743 [n_regions]"""
import gzip, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _harness import ref_side, read_fasta
from test_q4_theory_cpu import exact_colmax_and_carried_f

name = sys.argv[1] if len(sys.argv) > 1 else "H19"
n_regions = int(sys.argv[2]) if len(sys.argv) > 2 else 12
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
tot = hi = flagged = differ = 0
for reg in rng.sample(range(len(recs)), n_regions):
    seg = recs[reg][:5000]
    for para, strand, rule in TASKS:
        seq2, _ = S.task_strings(seg, para, strand, rule)
        exact, fmax = exact_colmax_and_carried_f(rna, seq2)
        tot += 1
        if exact.max() < 148 and fmax < 132:
            continue
        hi += 1
        if fmax < 132:
            continue
        flagged += 1
        got = S.colmax(rna, seq2)
        differ += int(not np.array_equal(exact, got))
print({"lncRNA": name, "nt": len(rna), "tasks": tot, "reach_148": hi, "flagged_by_carried_F": flagged, "really_different": differ})
