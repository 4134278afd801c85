"""CPU: the oracle (oracle/oracle.cpp) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

from _harness import GOLDEN, TASKS, oracle, oracle_side, params_array, read_fasta

O = oracle_side()


def _demo(data_dir):
    rna = read_fasta(os.path.join(data_dir, "H19.fa"))[0][1]
    hdr, dna = read_fasta(os.path.join(data_dir, "testDNA.fa"))[0]
    return rna, hdr, dna


def _tfosorted(rna, dna, chr_tag, start, **kw):
    cap = 1 << 24
    buf = C.create_string_buffer(cap)
    f = oracle().orc_run_tfosorted
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_long, C.POINTER(C.c_int), C.c_char_p, C.c_long]
    k = f(rna.encode(), dna.encode(), chr_tag.encode(), start, params_array(**kw), buf, cap)
    assert k >= 0
    return buf.value.decode()


def test_rule_translation_and_thresholds(golden, data_dir):
    rna, _, dna = _demo(data_dir)
    for t in golden["demo_tasks"][::4]:
        s2, src = O.task_strings(dna, t["para"], t["strand"], t["rule"])
        assert s2[:60] == t["seq2_head"] and src[:60] == t["src_head"]
        assert O.calc_score_once(rna, s2) == t["max_score"]


def test_colmax_and_peaks_demo(golden, data_dir):
    rna, _, dna = _demo(data_dir)
    cms = np.load(os.path.join(GOLDEN, "demo_colmax.npz"))["colmax"]
    for k in range(0, 48, 3):          # includes overflow (Q2) tasks: rule 10 / 16 anti-parallel
        t = golden["demo_tasks"][k]
        s2, _ = O.task_strings(dna, t["para"], t["strand"], t["rule"])
        cm = O.colmax(rna, s2)
        assert (cm == cms[k]).all()
        assert [list(p) for p in O.prealign(rna, s2, t["threshold"])] == t["peaks"]
    k = [i for i, t in enumerate(golden["demo_tasks"]) if (t["para"], t["strand"], t["rule"]) == (-1, 1, 10)][0]
    assert golden["demo_tasks"][k]["max_score"] >= 251 and golden["demo_tasks"][k]["peaks"] == []


def test_window_alignments_demo(golden, data_dir):
    rna, _, _ = _demo(data_dir)
    for w in golden["demo_windows"][::2]:
        o5, cig = O.align(rna, w["window"])
        assert list(o5) == w["out5"] and cig == w["cigar"]


def test_demo_tfosorted_bytes(data_dir):
    rna, hdr, dna = _demo(data_dir)
    sp, ch, rng = hdr.split("|")
    start = int(rng.split("-")[0])
    exp = open(os.path.join(GOLDEN, "demo_lg40__hg19-H19-testDNA-TFOsorted")).read()
    assert _tfosorted(rna, dna, ch, start, cLength=40) == exp
    assert len(exp.splitlines()) == 157


def test_demo_complex_flags_bytes(data_dir):
    rna, hdr, dna = _demo(data_dir)
    sp, ch, rng = hdr.split("|")
    start = int(rng.split("-")[0])
    exp = open(os.path.join(GOLDEN, "demo_complex__hg19-H19-testDNA-TFOsorted")).read()
    got = _tfosorted(rna, dna, ch, start, minIdentity=70, minStability=1, ntMin=25, ntMax=1000, penaltyC=1, penaltyT=-500,
                     cDistance=10, cLength=60)
    assert got == exp


def test_q4_reproducer(golden):
    q = golden["q4"]
    assert q["calc"] == 190 and max(q["colmax"]) == 181          # exact 190, reference's striped kernel 181
    assert list(O.colmax(q["rna"], q["dna"])) == q["colmax"]
    assert O.calc_score_once(q["rna"], q["dna"]) == q["calc"]
    o5, cig = O.align(q["rna"], q["dna"])
    assert [list(o5), cig] == q["align"]
    # the exact model differs from the literal kernel exactly here
    out = (C.c_int * len(q["dna"]))()
    tm = C.c_int(0)
    oracle().orc_colmax_model(q["rna"].encode(), q["dna"].encode(), len(q["dna"]), out, C.byref(tm))
    assert tm.value == 190 and list(out) != q["colmax"]


def test_n_and_u_scoring_split(golden):
    g = golden["nu"]
    assert O.calc_score_once(g["base"], g["dna_n"]) == g["calc_n"]
    assert list(O.colmax(g["base"], g["dna_n"])) == g["colmax_n"]
    assert O.calc_score_once(g["rna_u"], g["base"]) == g["calc_u"]
    assert list(O.colmax(g["rna_u"], g["base"])) == g["colmax_u"]
    assert g["calc_n"] != max(g["colmax_n"]) and g["calc_u"] != max(g["colmax_u"])


@pytest.mark.parametrize("key", ["tail", "homopolymer", "nu_record", "nu_record_plainrna"])
def test_record_level_cases(golden, key):
    g = golden[key]
    dna = g.get("dna", golden["nu_record"]["dna"])
    assert O.longtarget(g["rna"], dna, cLength=20) == g["text"]


def test_meg3_record(golden, data_dir):
    rna = read_fasta(os.path.join(data_dir, "MEG3-ENST00000451743.fa"))[0][1]
    recs = read_fasta(os.path.join(data_dir, "MEG3-DNAseq-first12.fa"))
    for k in (0, 7):
        hdr, dna = recs[k]
        sp, ch, rng = hdr.split("|")
        assert _tfosorted(rna, dna, ch, int(rng.split("-")[0]), cLength=60) == golden["meg3_records"]["rec%02d" % k]


def test_cluster_restatement():
    # hand-made triplexes: overlapping clusters, nt == length boundary (Q9), unassigned leftovers
    stari = [100, 104, 108, 300, 301, 500, 10, 700]
    endi = [160, 166, 170, 380, 381, 560, 70, 760]
    nt = [61, 63, 63, 81, 81, 61, 50, 61]
    mid, cen, mot = O.cluster(stari, endi, nt, 15, 50)
    assert mid[6] == 0 and mot[6] == 0            # nt == length is not counted
    assert mot[0] == mot[1] == mot[2] and len(set(mot) - {0}) >= 3


def test_oracle_lowercase_variant_equals_its_binary():
    """The oracle's restatement of the OLDER variant (fasim-LongTarget.cpp + fastSim.h: Params::lowercase) against files written
    by that variant's unmodified binary (tests/golden/make_golden_lowercase.py): demo, and 12 MEG3 regions as one multi-record run."""
    import os
    from _harness import GOLDEN, oracle_side, read_fasta
    O = oracle_side()
    data = os.path.join(GOLDEN, "data")
    rna = read_fasta(os.path.join(data, "H19.fa"))[0][1]
    hdr, dna = read_fasta(os.path.join(data, "testDNA.fa"))[0]
    sp, ch, rng = hdr.split("|")
    got, oob = O.run_lowercase(rna, [(dna, ch, int(rng.split("-")[0]))], cLength=40)
    assert got == open(os.path.join(GOLDEN, "demo_lc_lg40__hg19-H19-fastSim-TFOsorted")).read()
    rna = read_fasta(os.path.join(data, "MEG3-ENST00000451743.fa"))[0][1]
    recs = []
    for h, s in read_fasta(os.path.join(data, "MEG3-DNAseq-first12.fa")):
        sp, ch, rng = h.split("|")
        recs.append((s, ch, int(rng.split("-")[0])))
    got, oob = O.run_lowercase(rna, recs, cLength=60)
    assert got == open(os.path.join(GOLDEN, "meg3_first12_lc__MACS_pk13559-MEG3-ENST00000451743-fastSim-TFOsorted")).read()
