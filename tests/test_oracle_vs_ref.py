"""CPU: the oracle (oracle/oracle.cpp, the repo's CPU restatement) against the UNMODIFIED reference compiled from
/root/reference into oracle/_ref/libref_shim.so (oracle/Makefile, oracle/ref_shim.cpp) on seeded random inputs — this is
what pins the oracle beyond the stored golden vectors.  Skipped where the shim was not built (no reference sources)."""
import random

import pytest

from _harness import TASKS, have_ref_shim, oracle_side, ref_side, splitmix_bases

pytestmark = pytest.mark.skipif(not have_ref_shim(), reason="oracle/_ref/libref_shim.so not built (reference sources absent)")


def plant(dna, rna, at, k, rnd):
    """a noisy parallel-rule-1 target: RNA T->A, G->T read back through the rule gives a strong hit"""
    hit = list(rna[:k].translate(str.maketrans("TG", "AT")))
    for _ in range(max(1, k // 12)):
        hit[rnd.randrange(k)] = rnd.choice("ACGT")
    return dna[:at] + "".join(hit) + dna[at + k:]


@pytest.fixture(scope="module")
def sides():
    return oracle_side(), ref_side()


def test_translation_threshold_colmax_peaks(sides):
    O, R = sides
    rnd = random.Random(5)
    for trial in range(6):
        m, n = rnd.choice([40, 333, 800]), rnd.choice([200, 1500])
        rna, dna = splitmix_bases(300 + trial, m), splitmix_bases(400 + trial, n)
        if trial % 2 == 0:
            dna = plant(dna, rna, n // 3, min(m, 60), rnd)
        if trial == 3:
            dna = dna[:50] + "N" + dna[51:]
        for para, strand, rule in rnd.sample(TASKS, 6):
            so, sr = O.task_strings(dna, para, strand, rule), R.task_strings(dna, para, strand, rule)
            assert so == sr
            s2 = so[0]
            mo, mr = O.calc_score_once(rna, s2), R.calc_score_once(rna, s2)
            assert mo == mr
            assert (O.colmax(rna, s2) == R.colmax(rna, s2)).all()
            thr = int(mo * 0.8)
            assert O.prealign(rna, s2, thr) == R.prealign(rna, s2, thr)


def test_window_alignments(sides):
    O, R = sides
    rnd = random.Random(6)
    rna = splitmix_bases(77, 900)
    for trial in range(40):
        w = rnd.randrange(8, 160)
        win = splitmix_bases(500 + trial, w)
        if trial % 2 == 0:
            k = min(w, rnd.randrange(10, 70))
            at = rnd.randrange(0, 900 - k)
            seg = list(rna[at:at + k])
            if k > 20 and trial % 4 == 0:
                del seg[k // 2]                      # a gap
            win = win[:w - len(seg)] + "".join(seg)
        assert O.align(rna, win) == R.align(rna, win), trial


def test_task_and_record_level(sides):
    O, R = sides
    rnd = random.Random(7)
    rna = splitmix_bases(88, 600)
    dna = splitmix_bases(99, 5200)
    for at in (700, 2600, 4700):
        dna = plant(dna, rna[100:], at, 70, rnd)
    for para, strand, rule in [(1, 0, 1), (1, 1, 3), (-1, 1, 7), (-1, 0, 12)]:
        assert O.task(rna, dna[:5000], 0, para, strand, rule, cLength=20) == R.task(rna, dna[:5000], 0, para, strand, rule, cLength=20)
    for kw in (dict(cLength=20), dict(cLength=30, minIdentity=70, penaltyT=-500, penaltyC=1, ntMin=25), dict(cutLength=2000, overlap=300, cLength=20)):
        assert O.longtarget(rna, dna, **kw) == R.longtarget(rna, dna, **kw), kw
