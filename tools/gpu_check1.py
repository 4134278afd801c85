"""First on-GPU parity probe: scan stage (max / threshold / colmax / peaks), window alignments and the record-level
triplex list of the demo against the CPU oracle.  Writes a report to gpurun_out/check1.txt."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200"))
from _harness import *
import fasim_b200 as fb
import numpy as np

out = open(os.path.join(ROOT, "gpurun_out", "check1.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()

O = oracle_side()
D = os.path.join(ROOT, "tests", "golden", "data")
rna = read_fasta(os.path.join(D, "H19.fa"))[0][1]
hdr, dna = read_fasta(os.path.join(D, "testDNA.fa"))[0]
eng = fb.Engine(0)
eng.set_query("H19", rna)
t0 = time.time()
pr = eng.probe_segment(dna, TASKS)
P("probe time", time.time() - t0)
bad = 0
for (para, strand, rule), g in zip(TASKS, pr):
    s2, src = O.task_strings(dna, para, strand, rule)
    mx = O.calc_score_once(rna, s2)
    cm = O.colmax(rna, s2)
    pk = O.prealign(rna, s2, int(mx * 0.8))
    ok_m = mx == g["max_score"]; ok_c = (cm == g["colmax"]).all(); ok_p = [(s, p) for s, p in pk] == g["peaks"]
    if not (ok_m and ok_c and ok_p):
        bad += 1
        P("TASK", para, strand, rule, "max", mx, g["max_score"], "thr", int(mx*0.8), g["threshold"], "colmax_diff", int((cm != g["colmax"]).sum()),
          "peaks", len(pk), len(g["peaks"]), "literal", g["literal"])
        d = np.nonzero(cm != g["colmax"])[0][:5]
        P("   first diffs", [(int(j), int(cm[j]), int(g["colmax"][j])) for j in d])
P("scan-stage mismatching tasks:", bad, "of", len(TASKS), " literal tasks:", sum(g["literal"] for g in pr))

# windows: take oracle window schedule of 6 tasks
try:
    import ctypes as C
    wins = []
    for (para, strand, rule) in TASKS[:8]:
        s2, src = O.task_strings(dna, para, strand, rule)
        rows = (C.c_int * (8 * 4096))()
        k = oracle().orc_task_trace(rna.encode(), dna.encode(), para, strand, rule, params_array(), rows, 4096)
        for i in range(min(k, 4096)):
            ps, pp, cut, sw, rb, re_, qb, qe = rows[8*i:8*i+8]
            wins.append((s2[pp - cut + 1: pp + 1], (sw, rb, re_, qb, qe)))
    got = eng.Align([w for w, _ in wins])
    wb = 0
    for (w, exp), (g5, cig) in zip(wins, got):
        o5, ocig = O.align(rna, w)
        if tuple(o5) != tuple(g5) or [c for c in ocig if c >> 4] != cig:
            wb += 1
            if wb <= 10: P("WIN", len(w), "oracle", o5, "gpu", g5, "cig", ocig[:6], cig[:6])
    P("window mismatches:", wb, "of", len(wins))
except Exception:
    P(traceback.format_exc())

# record level
try:
    t0 = time.time()
    rows = eng.LongTarget(dna, "chr11", 2158478)
    P("LongTarget time", time.time() - t0, "rows", len(rows))
    txt = O.longtarget(rna, dna)
    exp = [l.split("\t") for l in txt.splitlines()]
    P("oracle rows", len(exp))
    import struct
    def f2b(x): return "%08x" % struct.unpack("<I", struct.pack("<f", x))[0]
    got = [[str(r["stari"]), str(r["endi"]), str(r["starj"]), str(r["endj"]), str(r["strand"]), str(r["reverse"]), str(r["rule"]), str(r["nt"]),
            f2b(r["score"]), f2b(r["identity"]), f2b(r["tri_score"]), r["tfo"], r["tts"]] for r in rows]
    exp2 = [e[:11] + [e[17], e[18]] for e in exp]
    P("record-level equal:", got == exp2)
    if got != exp2:
        for i, (a, b) in enumerate(zip(got, exp2)):
            if a != b: P("first diff row", i, a, b); break
except Exception:
    P(traceback.format_exc())
