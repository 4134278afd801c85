"""GPU parity tests: the CUDA path (through the C ABI of include/fasim_b200.h) against the golden vectors of the
unmodified reference and against the CPU oracle on the same seeded inputs.  Integer / index / byte results must be
bit-exact; the float32 columns (score, identity, stability) are compared by bit pattern as well (the stated
tolerance of 1e-6 relative is therefore met with margin 0)."""
import ctypes as C
import gzip
import os
import random
import struct
import subprocess
import sys

import numpy as np
import pytest

import fasim_b200 as fb
from _harness import GOLDEN, TASKS, have_ref_shim, oracle, oracle_side, params_array, read_fasta, ref_side, splitmix_bases

pytestmark = pytest.mark.gpu
O = oracle_side()


def f2b(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", x))[0]


def rows_as_oracle_text(rows):
    return [[str(r["stari"]), str(r["endi"]), str(r["starj"]), str(r["endj"]), str(r["strand"]), str(r["reverse"]), str(r["rule"]),
             str(r["nt"]), f2b(r["score"]), f2b(r["identity"]), f2b(r["tri_score"]), r["tfo"], r["tts"]] for r in rows]


def oracle_text_rows(txt):
    out = []
    for l in txt.splitlines():
        e = l.split("\t")
        out.append(e[:11] + [e[17], e[18]])
    return out


def demo(data_dir):
    rna = read_fasta(os.path.join(data_dir, "H19.fa"))[0][1]
    hdr, dna = read_fasta(os.path.join(data_dir, "testDNA.fa"))[0]
    return rna, hdr, dna


def run_cli_files(tmp_path, dna_name, dna_text, rna_name, rna_text, flags):
    d = str(tmp_path)
    open(os.path.join(d, dna_name), "w").write(dna_text)
    open(os.path.join(d, rna_name), "w").write(rna_text)
    os.makedirs(os.path.join(d, "out"), exist_ok=True)
    r = fb.run_cli(["-f1", dna_name, "-f2", rna_name, "-O", "out/"] + flags, cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    return {f: open(os.path.join(d, "out", f)).read() for f in sorted(os.listdir(os.path.join(d, "out")))}


# ------------------------------------------------------------------------------------------------ scan stage
def test_scan_stage_demo_all_tasks(engine, golden, data_dir):
    rna, _, dna = demo(data_dir)
    engine.set_params()
    engine.set_query("H19", rna)
    got = engine.probe_segment(dna, TASKS)
    cms = np.load(os.path.join(GOLDEN, "demo_colmax.npz"))["colmax"]
    for k, (t, g) in enumerate(zip(golden["demo_tasks"], got)):
        assert g["max_score"] == t["max_score"], t
        assert g["threshold"] == t["threshold"], t
        assert (g["colmax"] == cms[k]).all(), t
        assert [list(p) for p in g["peaks"]] == t["peaks"], t
    assert sum(1 for t in golden["demo_tasks"] if t["max_score"] >= 251) == 4      # Q2 tasks are in the fixture


def test_window_alignments_demo(engine, golden, data_dir):
    rna, _, _ = demo(data_dir)
    engine.set_params()
    engine.set_query("H19", rna)
    wins = golden["demo_windows"]
    got = engine.Align([w["window"] for w in wins])
    for w, (o5, cig) in zip(wins, got):
        assert list(o5) == w["out5"]
        assert cig == [c for c in w["cigar"] if c >> 4]          # zero-length ops of banded_sw are not observable


def test_q4_reproducer_window_seam(engine, golden):
    """Aligner::Align on the constructed Q4 case: the reference reports 181 / 30M2I12M where exact SW gives 190."""
    q = golden["q4"]
    engine.set_params()
    engine.set_query("q4", q["rna"])
    (o5, cig), = engine.Align([q["dna"]])
    assert [list(o5), cig] == [q["align"][0], [c for c in q["align"][1] if c >> 4]]
    assert o5[0] == 181


def test_q4_reproducer_scan_seam(engine, golden):
    """Column maxima of the Q4 case through the scan seam.  AntiMinus rule 8 is a bijection on ACGT
    (A->G, C->C, G->T, T->A), so the crafted translated text has a pre-image segment."""
    q = golden["q4"]
    seg = q["dna"].translate(str.maketrans("GCTA", "ACGT"))
    s2, _ = O.task_strings(seg, -1, 1, 8)
    assert s2 == q["dna"]
    engine.set_params()
    engine.set_query("q4", q["rna"])
    g = engine.probe_segment(seg, [(-1, 1, 8)])[0]
    assert g["literal"] == 1
    assert list(g["colmax"]) == q["colmax"]
    assert g["max_score"] == q["calc"] == 190 and max(q["colmax"]) == 181
    assert g["peaks"] == O.prealign(q["rna"], s2, int(190 * 0.8))


def test_n_and_u_scoring(engine, golden):
    g = golden["nu_record"]
    engine.set_params(c_length=20)
    engine.set_query("u", g["rna"])
    rows = engine.LongTarget(g["dna"])
    assert rows_as_oracle_text(rows) == oracle_text_rows(g["text"])
    g2 = golden["nu_record_plainrna"]
    engine.set_query("plain", g2["rna"])
    rows = engine.LongTarget(g["dna"])
    assert rows_as_oracle_text(rows) == oracle_text_rows(g2["text"])
    # function level: thresholds under the N-aware scoring, column maxima under the SSW scoring
    seg = g["dna"][900:1400]
    pr = engine.probe_segment(seg, TASKS[:6])
    for (pa, st, ru), r in zip(TASKS[:6], pr):
        s2, _ = O.task_strings(seg, pa, st, ru)
        assert r["max_score"] == O.calc_score_once(g2["rna"], s2)
        assert (r["colmax"] == O.colmax(g2["rna"], s2)).all()


@pytest.mark.parametrize("cut,overlap,n", [(6500, 0, 14000), (6500, 500, 9000), (300, 100, 3100), (150, 100, 1234), (400, 390, 700)])
def test_unusual_segmentations_vs_oracle(engine, cut, overlap, n):
    """-c / -o at the edges of the supported envelope: the largest cut length (16-bit cells), tiny cuts, stride 1."""
    rna = splitmix_bases(2001, 500)
    dna = list(splitmix_bases(1001, n))
    hit = rna[100:170].translate(str.maketrans("TG", "AT"))
    for at in range(200, n - 100, 1700):
        dna[at:at + 70] = hit
    dna = "".join(dna)
    engine.set_params(cut_length=cut, overlap=overlap, c_length=20)
    engine.set_query("r", rna)
    rows = engine.LongTarget(dna, "chrS", 1)
    assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna, cutLength=cut, overlap=overlap, cLength=20))
    assert len(rows) > 0


def test_cut_length_envelope(engine):
    """-c beyond 6500: fine while min(lncRNA, cut) keeps the scores inside 16 bits (the reference accepts any -c); a long lncRNA with a
    long cut, and cuts beyond the shared-memory limit, are refused loudly before any work."""
    rna = splitmix_bases(2001, 400)
    dna = list(splitmix_bases(1001, 21000))
    for at in (500, 9000, 15500, 20000):
        dna[at:at + 64] = rna[100:164].translate(str.maketrans("TG", "AT"))
    dna = "".join(dna)
    engine.set_params(cut_length=12000, overlap=1000, c_length=20)
    engine.set_query("short", rna)
    rows = engine.LongTarget(dna, "chrC", 1)
    assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna, cutLength=12000, overlap=1000, cLength=20)) and len(rows) > 3
    with pytest.raises(fb.FasimError):                       # 7000 nt x 12000 bp: scores could leave the 16-bit range
        engine.set_query("long", splitmix_bases(2002, 7000))
        engine.LongTarget(dna[:8000])
    with pytest.raises(fb.FasimError):
        engine.set_params(cut_length=30000)
    engine.set_params()
    engine.set_query("short", rna)


def test_microsatellite_stress_vs_oracle(engine):
    """Repeat-rich input: noisy (GA)n / (TC)n / (GGA)n stretches against an lncRNA with (CT)n / (GA)n / (GT)n tracts.  Hundreds
    of overlapping hits per task, run merging, >= 251 overflow (stop-recording), literal-scored tasks and windows, wide
    traceback bands, top-50 truncation — under strict, relaxed and positive-penalty filter settings."""
    rnd = random.Random(3)

    def noisy(unit, n, rate):
        s = list((unit * (n // len(unit) + 1))[:n])
        for i in range(n):
            if rnd.random() < rate:
                s[i] = rnd.choice("ACGT")
        return "".join(s)

    rna = (splitmix_bases(2001, 150) + noisy("CT", 130, 0.06) + splitmix_bases(2002, 100) + noisy("GA", 90, 0.04) + splitmix_bases(2003, 80)
           + noisy("GT", 70, 0.05) + splitmix_bases(2004, 60))
    dna = (splitmix_bases(1001, 1500) + noisy("GA", 1800, 0.08) + splitmix_bases(1002, 900) + noisy("TC", 700, 0.05) + splitmix_bases(1003, 1200)
           + noisy("GGA", 600, 0.1) + splitmix_bases(1004, 800))
    engine.set_query("stress", rna)
    total = 0
    for kw, ekw in ((dict(c_length=20), dict(cLength=20)),
                    (dict(c_length=15, nt_min=10, penalty_t=0, penalty_c=0, min_stability=0, min_identity=40),
                     dict(cLength=15, ntMin=10, penaltyT=0, penaltyC=0, minStability=0, minIdentity=40)),
                    (dict(c_length=30, penalty_t=-3, penalty_c=2, min_stability=1, min_identity=50, cut_length=3000, overlap=500),
                     dict(cLength=30, penaltyT=-3, penaltyC=2, minStability=1, minIdentity=50, cutLength=3000, overlap=500))):
        engine.set_params(**kw)
        res = engine.scan_record(dna, "chrM", 1)
        rows = fb.result_rows(res)
        lit = (res.contents.n_literal_tasks, res.contents.n_literal_windows, res.contents.n_q4_probed)
        engine.free(res)
        assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna, **ekw)), kw
        # the Q4 paths were exercised: flagged pairs went through the taint sweep (which may certify every one of them, so
        # literal TASKS are not guaranteed here; test_q4_probe_is_exact forces them), windows through the literal emulation
        assert lit[2] > 0 and (lit[1] > 0 or os.environ.get("LTG_WIN_Q4CHK") == "1")
        total += len(rows)
    assert total > 1000
    engine.set_params()


def test_multi_query_switching(engine):
    """Config 5 style use: several lncRNAs against the same DNA through one context, switching back and forth (different
    lengths select different scan tilings and strip counts; every buffer of the context is reused)."""
    dna = list(splitmix_bases(1001, 11000))
    rnas = [splitmix_bases(4001 + k, m) for k, m in enumerate([300, 1500, 777, 2100])]
    for k, r in enumerate(rnas):
        at = 900 + 2400 * k
        dna[at:at + 64] = r[20:84].translate(str.maketrans("TG", "AT"))
    dna = "".join(dna)
    engine.set_params(c_length=20)
    expect = [oracle_text_rows(O.longtarget(r, dna, cLength=20)) for r in rnas]
    for k in [0, 1, 2, 3, 1, 0, 3, 2]:
        engine.set_query("q%d" % k, rnas[k])
        assert rows_as_oracle_text(engine.LongTarget(dna, "chrQ", 1)) == expect[k], k
    assert all(len(e) > 0 for e in expect)


def test_n_in_some_segments_only(engine):
    """The N-aware threshold pass (Q3) runs only for the segments that contain a byte outside ACGT; the other segments of
    the same batch keep the single-pass threshold.  Three segments, N's in the middle one."""
    rna = splitmix_bases(2001, 600)
    dna = list(splitmix_bases(1001, 12000))
    hit = rna[200:260].translate(str.maketrans("TG", "AT"))
    for at in (1500, 6500, 11000):                       # the same planted target in every segment
        dna[at:at + 60] = hit
    for at in (6100, 6520, 6533, 7000):                  # N's: inside and next to the target of the second segment
        dna[at] = "N"
    dna = "".join(dna)
    engine.set_params(c_length=20)
    engine.set_query("r", rna)
    rows = engine.LongTarget(dna, "chrN", 1)
    assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna, cLength=20)) and len(rows) > 0


@pytest.mark.parametrize("key", ["tail", "homopolymer"])
def test_segment_edge_cases(engine, golden, key):
    g = golden[key]
    engine.set_params(c_length=20)
    engine.set_query("syn", g["rna"])
    rows = engine.LongTarget(g["dna"])
    assert rows_as_oracle_text(rows) == oracle_text_rows(g["text"])


def test_tiny_and_ragged_inputs_vs_oracle(engine):
    rnd = random.Random(11)
    engine.set_params(c_length=10, nt_min=5)
    for trial in range(12):
        m = rnd.choice([1, 2, 15, 16, 17, 31, 100, 511, 512, 513, 777])
        n = rnd.choice([1, 2, 5, 31, 32, 33, 100, 333])
        rna = splitmix_bases(100 + trial, m)
        dna = splitmix_bases(200 + trial, n)
        if trial % 3 == 0:            # plant a strong hit
            k = min(m, n, 60)
            dna = dna[:n - k] + rna[:k].replace("A", "x").replace("T", "A").replace("x", "T")[:k]
        engine.set_query("r", rna)
        pr = engine.probe_segment(dna, TASKS)
        for (pa, st, ru), r in zip(TASKS, pr):
            s2, _ = O.task_strings(dna, pa, st, ru)
            mx = O.calc_score_once(rna, s2)
            assert r["max_score"] == mx, (m, n, pa, st, ru)
            assert (r["colmax"] == O.colmax(rna, s2)).all(), (m, n, pa, st, ru)
            assert r["peaks"] == O.prealign(rna, s2, int(mx * 0.8)), (m, n, pa, st, ru)
        rows = engine.LongTarget(dna)
        assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna, cLength=10, ntMin=5)), (m, n)
    assert engine.LongTarget("") == []


# ------------------------------------------------------------------------------------------------ record / file level
def test_demo_record_level(engine, data_dir):
    rna, _, dna = demo(data_dir)
    engine.set_params()
    engine.set_query("H19", rna)
    rows = engine.LongTarget(dna, "chr11", 2158478)
    assert rows_as_oracle_text(rows) == oracle_text_rows(O.longtarget(rna, dna))
    assert all(r["genomestart"] == r["starj"] + 2158478 - 1 for r in rows)


@pytest.mark.parametrize("tag,flags", [
    ("lg40", ["-lg", "40"]),
    ("complex", ["-c", "5000", "-i", "70", "-S", "1.0", "-ni", "25", "-na", "1000", "-pc", "1", "-pt", "-500", "-ds", "10", "-lg", "60"]),
])
def test_cli_demo_files_byte_equal(tmp_path, data_dir, tag, flags):
    files = run_cli_files(tmp_path, "testDNA.fa", open(os.path.join(data_dir, "testDNA.fa")).read(), "H19.fa",
                          open(os.path.join(data_dir, "H19.fa")).read(), flags)
    names = [f for f in os.listdir(GOLDEN) if f.startswith("demo_%s__" % tag)]
    assert len(names) == 3
    for g in names:
        assert files[g.split("__", 1)[1]] == open(os.path.join(GOLDEN, g)).read(), g


@pytest.mark.parametrize("tag,flags", [("r3", ["-r", "3", "-lg", "40"]), ("t1", ["-t", "1", "-lg", "40"]),
                                       ("tm1r10", ["-t", "-1", "-r", "10", "-lg", "30"]), ("c1000", ["-c", "1000", "-o", "50", "-lg", "40"])])
def test_cli_rule_strand_cut_options(tmp_path, data_dir, tag, flags):
    files = run_cli_files(tmp_path, "testDNA.fa", open(os.path.join(data_dir, "testDNA.fa")).read(), "H19.fa",
                          open(os.path.join(data_dir, "H19.fa")).read(), flags)
    assert files["hg19-H19-testDNA-TFOsorted"] == open(os.path.join(GOLDEN, "demo_%s__TFOsorted" % tag)).read()


def test_cli_meg3_multi_record(tmp_path, data_dir, golden):
    files = run_cli_files(tmp_path, "MEG3-12.fa", open(os.path.join(data_dir, "MEG3-DNAseq-first12.fa")).read(), "MEG3.fa",
                          open(os.path.join(data_dir, "MEG3-ENST00000451743.fa")).read(), ["-lg", "60"])
    got = [v for k, v in files.items() if k.endswith("TFOsorted")][0]
    assert got == open(os.path.join(GOLDEN, "meg3_first12_mr__TFOsorted")).read()


def test_cli_meg3_all_532_regions_default_flags(tmp_path, data_dir):
    """BASELINE.json configs[1]: the MEG3 lncRNA against its 532 example regions, default flags (-c 5000 -i 60).  Golden =
    the reference run on the whole file (tests/golden/make_golden_meg3_full.py)."""
    dna = gzip.open(os.path.join(data_dir, "MEG3-DNAseq.fa.gz")).read().decode()
    files = run_cli_files(tmp_path, "MEG3-DNAseq.fa", dna, "MEG3.fa", open(os.path.join(data_dir, "MEG3-ENST00000451743.fa")).read(), [])
    got = [v for k, v in files.items() if k.endswith("TFOsorted")][0]
    exp = gzip.open(os.path.join(GOLDEN, "meg3_full_mr_defaults__TFOsorted.gz")).read().decode()
    assert len(exp.splitlines()) > 3000
    assert got == exp


def test_scan_records_equals_record_by_record(engine, data_dir):
    """ltg_scan_records (all records share the device batches) == ltg_scan_record per record, rows and order."""
    rna = read_fasta(os.path.join(data_dir, "MEG3-ENST00000451743.fa"))[0][1]
    recs = read_fasta(os.path.join(data_dir, "MEG3-DNAseq-first12.fa"))
    engine.set_params(c_length=30)
    engine.set_query("MEG3", rna)
    one_by_one, batch_in = [], []
    for k, (hdr, dna) in enumerate(recs):
        sp, ch, rng = hdr.split("|")
        start = int(rng.split("-")[0])
        rows = engine.LongTarget(dna, ch, start)
        for r in rows:
            r["record"] = k
        one_by_one += rows
        batch_in.append((dna, ch, start))
    res = engine.scan_records(batch_in + [("", "chrE", 5)])        # an empty record in the mix
    got = fb.result_rows(res)
    engine.free(res)
    assert got == one_by_one and len(got) > 50


def test_meg3_single_records(engine, data_dir, golden):
    rna = read_fasta(os.path.join(data_dir, "MEG3-ENST00000451743.fa"))[0][1]
    recs = read_fasta(os.path.join(data_dir, "MEG3-DNAseq-first12.fa"))
    engine.set_params(c_length=60)
    engine.set_query("MEG3", rna)
    for k, (hdr, dna) in enumerate(recs):
        sp, ch, rng = hdr.split("|")
        res = engine.scan_record(dna, ch, int(rng.split("-")[0]))
        engine.cluster_triplex(res)
        path = "/tmp/_meg3_rec_TFOsorted"
        engine.printResult(res, path)
        engine.free(res)
        assert open(path).read() == golden["meg3_records"]["rec%02d" % k], k


@pytest.mark.parametrize("name", ["NEAT1", "MALAT1"])
def test_cli_long_lncrnas_complex_flags(tmp_path, data_dir, name):
    files = run_cli_files(tmp_path, "testDNA.fa", open(os.path.join(data_dir, "testDNA.fa")).read(), name + ".fa",
                          open(os.path.join(data_dir, name + ".fa")).read(),
                          ["-i", "70", "-S", "1.0", "-ni", "25", "-pt", "-500", "-ds", "10", "-lg", "60"])
    got = [v for k, v in files.items() if k.endswith("TFOsorted")][0]
    assert got == open(os.path.join(GOLDEN, "%s_testDNA_complex__TFOsorted" % name)).read()


@pytest.mark.parametrize("name,frec,winchk", [("NEAT1", None, None), ("MALAT1", None, None), ("NEAT1", "1", None), ("MALAT1", "1", "0"),
                                              ("MALAT1", "0", "1"), ("NEAT1", None, "1"), ("NEAT1", None, "0")])
def test_cli_long_lncrnas_vs_meg3_regions(tmp_path, data_dir, name, frec, winchk, monkeypatch):
    """BASELINE configs[2] on the second substitute DNA set: NEAT1 (22.8 knt) / MALAT1 (8.7 knt) against the first 12 MEG3
    regions (multi-record file), complex flags; golden from the reference (tests/golden/make_golden_config3.py).  LTG_FREC: the
    Q4 verdict from the carried-F blocks the main sweep records (1), from the probe sweep (0), or chosen per query (unset)."""
    if frec is not None:
        monkeypatch.setenv("LTG_FREC", frec)
    if winchk is not None:               # window sweeps that watch the stripe starts (only flagged windows are emulated literally): forced / off
        monkeypatch.setenv("LTG_WIN_Q4CHK", winchk)
    files = run_cli_files(tmp_path, "MEG3-12.fa", open(os.path.join(data_dir, "MEG3-DNAseq-first12.fa")).read(), name + ".fa",
                          open(os.path.join(data_dir, name + ".fa")).read(),
                          ["-i", "70", "-S", "1.0", "-ni", "25", "-pt", "-500", "-ds", "10", "-lg", "60"])
    got = [v for k, v in files.items() if k.endswith("TFOsorted")][0]
    assert got == open(os.path.join(GOLDEN, "%s_meg3first12_complex__TFOsorted" % name)).read()


def test_cli_multi_device_work_queue(tmp_path):
    """--devices: one context + host thread per entry pulling shards / record groups from a shared queue.  Two contexts on
    GPU 0 ("0,0") exercise the path on a single-GPU box; the files must equal the single-context run byte for byte."""
    rna = splitmix_bases(2001, 1200)
    recs = [(11_000_000, 1001), (400_000, 1002), (2_600_000, 1003), (37, 1004), (5_000_000, 1005), (900_000, 1006)]
    fasta = "".join(">syn|chr%d|%d-%d\n%s\n" % (k + 1, 1000 * k + 1, 1000 * k + n, splitmix_bases(seed, n)) for k, (n, seed) in enumerate(recs))
    d = str(tmp_path)
    open(os.path.join(d, "multi.fa"), "w").write(fasta)
    open(os.path.join(d, "rna.fa"), "w").write(">synRNA\n%s\n" % rna)
    outs = []
    for tag, extra in (("one", ["--device", "0"]), ("two", ["--devices", "0,0"])):
        os.makedirs(os.path.join(d, tag), exist_ok=True)
        r = fb.run_cli(["-f1", "multi.fa", "-f2", "rna.fa", "-O", tag + "/", "-lg", "30"] + extra, cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append({f: open(os.path.join(d, tag, f)).read() for f in sorted(os.listdir(os.path.join(d, tag)))})
    assert outs[0] == outs[1] and len(outs[0]) == 3
    assert len([v for k, v in outs[0].items() if k.endswith("TFOsorted")][0].splitlines()) > 1000


def test_cli_multi_query_work_queue(tmp_path):
    """--queries (the multi-query work-queue path): a multi-record -f2 file; the (lncRNA, chunk) jobs go
    through the shared queue of two contexts.  Every lncRNA's files must equal those of a single-query run of the plain CLI."""
    lens = [1000 + 977 * k for k in range(4)] + [312]
    rnas = [("synRNA%d" % k, splitmix_bases(4001 + k, m)) for k, m in enumerate(lens)]
    recs = [(11_000_000, 1002), (600_000, 1003), (41, 1004)]
    planted = list(splitmix_bases(1002, recs[0][0]))
    for k, (_, r) in enumerate(rnas):                     # one planted target per lncRNA so that every output has rows
        at = 300_000 + 1_700_000 * k
        planted[at:at + 80] = r[100:180].translate(str.maketrans("TG", "AT"))
    seqs = ["".join(planted)] + [splitmix_bases(seed, n) for n, seed in recs[1:]]
    d = str(tmp_path)
    open(os.path.join(d, "dna.fa"), "w").write("".join(">syn|chr%d|%d-%d\n%s\n" % (k + 1, 1, len(sq), sq) for k, sq in enumerate(seqs)))
    open(os.path.join(d, "all.fa"), "w").write("".join(">%s\n%s\n" % (n, "\n".join(r[i:i + 70] for i in range(0, len(r), 70))) for n, r in rnas))
    os.makedirs(os.path.join(d, "multi"))
    r = fb.run_cli(["-f1", "dna.fa", "-f2", "all.fa", "-O", "multi/", "-lg", "30", "--queries", "--devices", "0,0"], cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    multi = {f: open(os.path.join(d, "multi", f)).read() for f in sorted(os.listdir(os.path.join(d, "multi")))}
    assert len(multi) == 3 * len(rnas)
    for name, rna in rnas:
        open(os.path.join(d, name + ".fa"), "w").write(">%s\n%s\n" % (name, rna))
        os.makedirs(os.path.join(d, name))
        r = fb.run_cli(["-f1", "dna.fa", "-f2", name + ".fa", "-O", name + "/", "-lg", "30"], cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
        for f in os.listdir(os.path.join(d, name)):
            assert multi[f] == open(os.path.join(d, name, f)).read(), f
            if f.endswith("TFOsorted"):
                assert len(multi[f].splitlines()) > 1, f


def test_cli_gzip_and_twobit_inputs(tmp_path):
    """-f1 as gzip FASTA and as UCSC .2bit (+ --seq regions): the output files equal those of the plain FASTA of the same bases."""
    import gzip
    from test_inputs_cpu import write_twobit
    rna = splitmix_bases(2001, 900)
    chrom = list(splitmix_bases(1001, 30000))
    for at in (4000, 12000, 21000):
        chrom[at:at + 70] = rna[100:170].translate(str.maketrans("TG", "AT"))
    chrom[15000:15040] = "N" * 40
    chrom = "".join(chrom)
    d = str(tmp_path)
    open(os.path.join(d, "rna.fa"), "w").write(">q\n%s\n" % rna)
    lo, hi = 2001, 26000                                        # 1-based inclusive region
    fasta = ">sp|chrT|%d-%d\n%s\n" % (lo, hi, chrom[lo - 1:hi])
    open(os.path.join(d, "reg.fa"), "w").write(fasta)
    with gzip.open(os.path.join(d, "reg.fa.gz"), "wt") as f:
        f.write(fasta)
    write_twobit(os.path.join(d, "reg.2bit"), [("chrOther", "ACGT" * 10), ("chrT", chrom)])
    outs = {}
    for tag, args in (("plain", ["-f1", "reg.fa"]), ("gz", ["-f1", "reg.fa.gz"]),
                      ("twobit", ["-f1", "reg.2bit", "--seq", "chrT:%d-%d" % (lo, hi), "--species", "sp"])):
        os.makedirs(os.path.join(d, tag))
        r = fb.run_cli(args + ["-f2", "rna.fa", "-O", tag + "/", "-lg", "30"], cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
        names = sorted(os.listdir(os.path.join(d, tag)))
        assert names == ["sp-q-reg-TFOclass1-15-30", "sp-q-reg-TFOclass2-15-30", "sp-q-reg-TFOsorted"], names
        outs[tag] = [open(os.path.join(d, tag, f)).read() for f in names]
    assert outs["gz"] == outs["plain"] and outs["twobit"] == outs["plain"]
    assert len(outs["plain"][2].splitlines()) > 3


def test_cli_synthetic_and_planted(tmp_path, golden):
    sdna, srna = splitmix_bases(1001, 30000), splitmix_bases(2001, 1000)
    files = run_cli_files(tmp_path, "syn.fa", ">syn|chr1|1-30000\n%s\n" % sdna, "synRNA.fa", ">synRNA1k\n%s\n" % srna, ["-lg", "20"])
    assert files["syn-synRNA1k-syn-TFOsorted"] == open(os.path.join(GOLDEN, "syn30k_lg20__TFOsorted")).read()
    p = golden["planted"]
    files = run_cli_files(tmp_path, "pl.fa", ">syn|chr1|1-20000\n%s\n" % p["dna"], "plRNA.fa", ">plRNA\n%s\n" % p["rna"], ["-lg", "30"])
    assert files["syn-plRNA-pl-TFOsorted"] == open(os.path.join(GOLDEN, "planted20k_lg30__TFOsorted")).read()


def test_reference_binary_repeat_rich_chunks(tmp_path):
    """At-scale file parity against the UNMODIFIED reference binary run on the box's host cores (oracle/_ref travels with the
    repo): 16 chunks of 60 kb of repeat-rich DNA - microsatellites with substitutions, purine tracts, one Alu-like element
    copied with mutations into every chunk, N runs - against an lncRNA with (CT)/(GA)/(GT)-rich tracts (overflows >= 251,
    long hit runs, top-50 truncation, literal-emulated tasks).  All three output files of every chunk must be byte-equal."""
    import random
    import subprocess
    from _harness import ref_binary
    if not os.path.exists(ref_binary()):
        pytest.skip("oracle/_ref/fasim not built")
    rng = random.Random(20240607)
    def mutate(s, rate):
        return "".join(rng.choice("ACGT") if rng.random() < rate else c for c in s)
    rna = list(splitmix_bases(2001, 1500))
    for at, unit in ((200, "CT"), (700, "GA"), (1200, "GT")):
        rna[at:at + 160] = mutate(unit * 80, 0.08)
    rna = "".join(rna)
    alu = splitmix_bases(555, 150) + mutate("GA" * 40, 0.05) + splitmix_bases(556, 70)
    d = str(tmp_path)
    open(os.path.join(d, "rna.fa"), "w").write(">lnc\n%s\n" % rna)
    n_chunks, n = 16, 60000
    for k in range(n_chunks):
        dna = list(splitmix_bases(9000 + k, n))
        for _ in range(15):
            at = rng.randrange(0, n - 400)
            unit = rng.choice(["GA", "CT", "GAA", "CCT", "A", "GGA", "TC", "AG"])
            L = rng.randrange(30, 260)
            dna[at:at + L] = mutate((unit * L)[:L], rng.choice([0.0, 0.05, 0.12]))
        for _ in range(5):
            at = rng.randrange(0, n - 400)
            dna[at:at + len(alu)] = mutate(alu, 0.06)
        at = rng.randrange(0, n - 100)
        L = rng.randrange(1, 40)
        dna[at:at + L] = "N" * L
        dna = "".join(dna)[:n]
        open(os.path.join(d, "c%02d.fa" % k), "w").write(">syn|chr%d|%d-%d\n%s\n" % (k + 1, 1000 * k + 1, 1000 * k + len(dna), dna))
    os.makedirs(os.path.join(d, "ref"))
    os.makedirs(os.path.join(d, "gpu"))
    procs = [subprocess.Popen([ref_binary(), "-f1", "c%02d.fa" % k, "-f2", "rna.fa", "-O", "ref/", "-lg", "30"], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for k in range(n_chunks)]
    for k in range(n_chunks):
        r = fb.run_cli(["-f1", "c%02d.fa" % k, "-f2", "rna.fa", "-O", "gpu/", "-lg", "30"], cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
    for pr in procs:
        assert pr.wait(timeout=900) == 0
    names = sorted(os.listdir(os.path.join(d, "ref")))
    assert len(names) == 3 * n_chunks and names == sorted(os.listdir(os.path.join(d, "gpu")))
    rows = 0
    for f in names:
        want, got = open(os.path.join(d, "ref", f)).read(), open(os.path.join(d, "gpu", f)).read()
        assert got == want, f
        rows += len(want.splitlines()) - 1 if f.endswith("TFOsorted") else 0
    assert rows > 1500


def _run_reference_and_cli(d, rna_files, chunk_files, flags, extra_cli=()):
    """Runs the unmodified reference binary (host cores, all (lncRNA, chunk) pairs in parallel) and this build's CLI on the
    same files; returns ({name: text} reference, {name: text} this build).  One reference process per (lncRNA, chunk)."""
    from _harness import ref_binary
    os.makedirs(os.path.join(d, "ref"))
    os.makedirs(os.path.join(d, "gpu"))
    procs = [subprocess.Popen([ref_binary(), "-f1", c, "-f2", r, "-O", "ref/"] + flags, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
             for r in rna_files for c in chunk_files]
    for r in rna_files:
        for c in chunk_files:
            out = fb.run_cli(["-f1", c, "-f2", r, "-O", "gpu/"] + flags + list(extra_cli), cwd=d)
            assert out.returncode == 0, out.stdout + out.stderr
    for pr in procs:
        assert pr.wait(timeout=1500) == 0
    rd = lambda sub: {f: open(os.path.join(d, sub, f)).read() for f in sorted(os.listdir(os.path.join(d, sub)))}
    return rd("ref"), rd("gpu")


def test_reference_binary_planted_bench_chunks(tmp_path):
    """The PLANTED variant of the headline workload: the bench's 100 Mbp region (seed 1001) with 160 bp of
    (GA)n carrying 10 % substitutions (SplitMix64 stream 3001) written over offset 500 000 of every Mbp, against the bench's
    3 kb lncRNA with three 200-nt (CT) / (GA) / (GT) tracts at 500 / 1500 / 2500 -> overflows >= 251, long runs of hits, top-50
    truncation.  Six 250 kb chunks centred on planted sites, every output file byte-equal to the unmodified reference."""
    from _harness import ref_binary
    if not os.path.exists(ref_binary()):
        pytest.skip("oracle/_ref/fasim not built")
    import bench
    rna = list(splitmix_bases(2001, 3000))
    for at, unit in ((500, "CT"), (1500, "GA"), (2500, "GT")):
        rna[at:at + 200] = unit * 100
    rna = "".join(rna)
    sub = splitmix_bases(3001, 160 * 100)
    d = str(tmp_path)
    open(os.path.join(d, "rna.fa"), "w").write(">synRNA3kP\n%s\n" % rna)
    chunk_files = []
    for k, mbp in enumerate((0, 17, 38, 55, 76, 99)):
        lo = mbp * 1_000_000 + 375_000
        dna = list(bench.splitmix_bases(1001, 250_000, lo).tobytes().decode())
        tract = list("GA" * 80)
        for i in range(160):
            z = sub[mbp * 160 + i]
            if (ord(z) * 7 + i) % 10 == 0:                   # ~10 % substitutions, deterministic
                tract[i] = z
        at = mbp * 1_000_000 + 500_000 - lo
        dna[at:at + 160] = tract
        name = "p%02d.fa" % k
        open(os.path.join(d, name), "w").write(">syn|chr1|%d-%d\n%s\n" % (lo + 1, lo + 250_000, "".join(dna)))
        chunk_files.append(name)
    ref, got = _run_reference_and_cli(d, ["rna.fa"], chunk_files, [])
    assert sorted(ref) == sorted(got) and len(ref) == 3 * len(chunk_files)
    rows = 0
    for f in ref:
        assert got[f] == ref[f], f
        rows += len(ref[f].splitlines()) - 1 if f.endswith("TFOsorted") else 0
    assert rows > 300


def test_reference_binary_multi_query_inputs(tmp_path):
    """BASELINE.json configs[4] inputs (DNA seed 1002, lncRNAs of 1000 + z % 9001 nt from seeds 4001..4004)
    through the multi-query work-queue path (`--queries --devices 0,0`): every lncRNA's files equal the unmodified reference
    run on that lncRNA alone (one reference process per lncRNA; the reference reads one lncRNA per run)."""
    from _harness import ref_binary
    if not os.path.exists(ref_binary()):
        pytest.skip("oracle/_ref/fasim not built")
    import bench
    qs = bench.synthetic_queries(4)
    dna = list(bench.splitmix_bases(1002, 120_000, 7 * 4900).tobytes().decode())
    for k, (_, r) in enumerate(qs):                        # one planted target per lncRNA so that every output has rows
        at = 9_000 + 27_000 * k
        dna[at:at + 90] = r[200:290].translate(str.maketrans("TG", "AT"))
    d = str(tmp_path)
    open(os.path.join(d, "dna.fa"), "w").write(">syn|chr1|%d-%d\n%s\n" % (7 * 4900 + 1, 7 * 4900 + len(dna), "".join(dna)))
    open(os.path.join(d, "all.fa"), "w").write("".join(">%s\n%s\n" % (n, r) for n, r in qs))
    os.makedirs(os.path.join(d, "ref"))
    os.makedirs(os.path.join(d, "gpu"))
    procs = []
    for n, r in qs:
        open(os.path.join(d, n + ".fa"), "w").write(">%s\n%s\n" % (n, r))
        procs.append(subprocess.Popen([ref_binary(), "-f1", "dna.fa", "-f2", n + ".fa", "-O", "ref/", "-lg", "30"], cwd=d,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    out = fb.run_cli(["-f1", "dna.fa", "-f2", "all.fa", "-O", "gpu/", "-lg", "30", "--queries", "--devices", "0,0"], cwd=d)
    assert out.returncode == 0, out.stdout + out.stderr
    for pr in procs:
        assert pr.wait(timeout=1500) == 0
    names = sorted(os.listdir(os.path.join(d, "ref")))
    assert len(names) == 3 * len(qs) and names == sorted(os.listdir(os.path.join(d, "gpu")))
    for f in names:
        assert open(os.path.join(d, "gpu", f)).read() == open(os.path.join(d, "ref", f)).read(), f
        if f.endswith("TFOsorted"):
            assert len(open(os.path.join(d, "ref", f)).read().splitlines()) > 1, f


def test_scan_packed_equals_text(engine):
    """ltg_scan_packed: DNA kept 2-bit packed (host bytes, or a packed store resident in HBM), expanded and cut into segments on
    the device.  Same rows as the text path, for a region that starts inside a byte, with N runs, as whole record and as shards."""
    import torch
    rna = splitmix_bases(2001, 800)
    chrom = list(splitmix_bases(1001, 26000))
    for at in (3000, 9000, 17000, 23000):
        chrom[at:at + 70] = rna[100:170].translate(str.maketrans("TG", "AT"))
    nruns = [(5000, 37), (12001, 1), (20000, 300)]
    for a, n in nruns:
        chrom[a:a + n] = "N" * n
    chrom = "".join(chrom)
    code = {"T": 0, "C": 1, "A": 2, "G": 3, "N": 0}
    packed = bytearray((len(chrom) + 3) // 4)
    for i, c in enumerate(chrom):
        packed[i // 4] |= code[c] << (6 - 2 * (i % 4))
    lo, hi = 1003, 25001                                        # region [lo, hi): starts 3 bases into a byte
    blocks = [(max(a, lo) - lo, min(a + n, hi) - max(a, lo)) for a, n in nruns if min(a + n, hi) > max(a, lo)]
    engine.set_params(c_length=20)
    engine.set_query("q", rna)
    res = engine.scan_record(chrom[lo:hi], "chrP", lo + 1)
    want, h2d_text = fb.result_rows(res), res.contents.h2d_bytes
    engine.free(res)
    assert len(want) > 3
    res = engine.scan_packed(bytes(packed), lo, hi - lo, blocks, "chrP", lo + 1)
    got = fb.result_rows(res)
    h2d = res.contents.h2d_bytes
    engine.free(res)
    assert got == want
    assert h2d_text - h2d > 0.7 * (hi - lo)                      # a quarter byte per base crossed PCIe, not one
    dev = torch.frombuffer(packed, dtype=torch.uint8).cuda()     # the packed store resident in HBM
    h2d_shards = {}
    for where in ("host", "device"):
        parts, total = [], 0
        for r in range(3):
            first, count, first_byte, n_bytes = fb.shard_segments(hi - lo, 3, r)
            sb = [(max(a, first_byte) - first_byte, min(a + n, first_byte + n_bytes) - max(a, first_byte)) for a, n in blocks
                  if min(a + n, first_byte + n_bytes) > max(a, first_byte)]
            res = engine.scan_packed(bytes(packed) if where == "host" else None, lo + first_byte, n_bytes, sb, "chrP", lo + 1, record_len=hi - lo,
                                     first_segment=first, n_segments=count, device_ptr=dev.data_ptr() if where == "device" else None)
            parts.append(fb.result_rows(res))
            total += res.contents.h2d_bytes
            engine.free(res)
        assert fb.merge_shard_rows(parts) == want
        h2d_shards[where] = total
    # with the packed store resident in HBM only descriptors, job lists and string jobs cross PCIe: the same shards fed from host
    # memory copy their packed bytes (a quarter byte per base, the shards overlap by a segment tail) on top of exactly that
    saved = h2d_shards["host"] - h2d_shards["device"]
    assert (hi - lo) // 4 <= saved <= (hi - lo) // 4 + 3 * (5000 // 4 + 8)
    engine.set_params()


# ------------------------------------------------------------------------------------------------ --compat lowercase
def test_cli_compat_lowercase_byte_equal(tmp_path, data_dir):
    """`fasim --compat lowercase`: the older variant that ships next to the canonical one (fasim-LongTarget.cpp + fastSim.h;
    window loop without start clamp, acceptance on equality, no per-task filter, -fastSim-TFOsorted naming).  Byte-equal to the
    UNMODIFIED older binary (tests/golden/make_golden_lowercase.py) on the demo and on the first 12 MEG3 regions (multi-record)."""
    files = run_cli_files(tmp_path, "testDNA.fa", open(os.path.join(data_dir, "testDNA.fa")).read(), "H19.fa",
                          open(os.path.join(data_dir, "H19.fa")).read(), ["-lg", "40", "--compat", "lowercase"])
    assert list(files) == ["hg19-H19-fastSim-TFOsorted"]
    assert files["hg19-H19-fastSim-TFOsorted"] == open(os.path.join(GOLDEN, "demo_lc_lg40__hg19-H19-fastSim-TFOsorted")).read()
    d2 = tmp_path / "m"
    d2.mkdir()
    files = run_cli_files(d2, "MEG3-12.fa", open(os.path.join(data_dir, "MEG3-DNAseq-first12.fa")).read(), "MEG3.fa",
                          open(os.path.join(data_dir, "MEG3-ENST00000451743.fa")).read(), ["-lg", "60", "--compat", "lowercase"])
    (name, text), = files.items()
    assert text == open(os.path.join(GOLDEN, "meg3_first12_lc__" + name)).read() and len(text.splitlines()) > 30


def test_compat_lowercase_vs_oracle(engine):
    """The same mode through the C ABI against the oracle's restatement of the older variant (pinned to its binary by
    tests/test_oracle_golden.py): planted targets right at segment starts, where the unclamped window offset matters."""
    rna = splitmix_bases(2001, 700)
    dna = list(splitmix_bases(1001, 9000))
    for at, (a, L) in zip((3, 40, 1500, 2980, 3010, 5990, 8950), ((50, 60), (200, 45), (300, 70), (400, 66), (90, 50), (500, 58), (610, 40))):
        dna[at:at + L] = rna[a:a + L].translate(str.maketrans("TG", "AT"))
    dna = "".join(dna)
    try:
        engine.set_compat(True)
        for kw, ekw in ((dict(c_length=20, cut_length=3000, overlap=100), dict(cLength=20, cutLength=3000, overlap=100)),
                        (dict(c_length=15, nt_min=30, min_identity=40, min_stability=0, penalty_t=-2), dict(cLength=15, ntMin=30, minIdentity=40, minStability=0, penaltyT=-2))):
            engine.set_params(**kw)
            engine.set_query("r", rna)
            res = engine.scan_record(dna, "chrL", 5)
            engine.cluster_triplex(res)
            engine.printResult(res, "/tmp/_lc_TFOsorted")
            engine.free(res)
            want, oob = O.run_lowercase(rna, [(dna, "chrL", 5)], **ekw)
            assert open("/tmp/_lc_TFOsorted").read() == want, kw
            assert len(want.splitlines()) > 5
    finally:
        engine.set_compat(False)
        engine.set_params()


# ------------------------------------------------------------------------------------------------ -F mode (SIM)
def test_sim_mode_vs_oracle(engine):
    """-F: every task through SIM() (sim.h:410).  k_sim's wavefront first pass + in-order node-list replay + the shared k-best
    core against the oracle restatement (itself pinned to the unmodified sim.h by tests/test_sim_cpu.py): planted targets on
    random DNA (several segments incl. a tail), and a repeat-rich record where the 50-node list overflows."""
    rnd = random.Random(23)
    try:
        engine.set_sim_mode(True)
        rna = splitmix_bases(2001, 333)
        dna = list(splitmix_bases(1001, 2300))
        for at in (150, 900, 1400, 2100):
            L = rnd.randrange(30, 70)
            a = rnd.randrange(0, len(rna) - L)
            dna[at:at + L] = rna[a:a + L].translate(str.maketrans("TG", "AT"))
        dna = "".join(dna)
        for kw, ekw in ((dict(c_length=20, cut_length=1000, overlap=100), dict(cLength=20, cutLength=1000, overlap=100)),
                        (dict(c_length=25, nt_min=22, nt_max=60, penalty_t=-3, penalty_c=2, cut_length=2500, overlap=50),
                         dict(cLength=25, ntMin=22, ntMax=60, penaltyT=-3, penaltyC=2, cutLength=2500, overlap=50))):
            engine.set_params(**kw)
            engine.set_query("r", rna)
            rows = engine.LongTarget(dna, "chrS", 1)
            assert rows_as_oracle_text(rows) == oracle_text_rows(O.sim_longtarget(rna, dna, **ekw)), kw
            assert len(rows) > 5

        def noisy(unit, n, rate):
            s = list((unit * (n // len(unit) + 1))[:n])
            for i in range(n):
                if rnd.random() < rate:
                    s[i] = rnd.choice("ACGT")
            return "".join(s)
        rna = splitmix_bases(2001, 60) + noisy("CT", 70, 0.08) + splitmix_bases(2002, 40) + noisy("GA", 50, 0.05)
        dna = splitmix_bases(1001, 150) + noisy("GA", 260, 0.1) + splitmix_bases(1002, 100) + noisy("TC", 150, 0.06)
        engine.set_params(c_length=15, nt_min=10, cut_length=400, overlap=60)
        engine.set_query("rr", rna)
        rows = engine.LongTarget(dna, "chrR", 1)
        assert rows_as_oracle_text(rows) == oracle_text_rows(O.sim_longtarget(rna, dna, cLength=15, ntMin=10, cutLength=400, overlap=60))
        assert len(rows) > 100
        # a 35-row lncRNA (one partial strip), a 1-column segment, an empty record
        engine.set_params(c_length=10, nt_min=5)
        engine.set_query("tiny", splitmix_bases(7, 35))
        for d in (splitmix_bases(8, 1), splitmix_bases(9, 90), ""):
            assert rows_as_oracle_text(engine.LongTarget(d)) == oracle_text_rows(O.sim_longtarget(splitmix_bases(7, 35), d, cLength=10, ntMin=5)) if d else engine.LongTarget(d) == []
    finally:
        engine.set_sim_mode(False)
        engine.set_params()


def test_cli_sim_mode_demo_byte_equal(tmp_path, data_dir):
    """`fasim -F` on the H19 / testDNA demo (-lg 40): byte-equal to the unmodified reference's -F run (120 s on a CPU core;
    golden generated by tests/golden/make_golden.py --sim)."""
    files = run_cli_files(tmp_path, "testDNA.fa", open(os.path.join(data_dir, "testDNA.fa")).read(), "H19.fa",
                          open(os.path.join(data_dir, "H19.fa")).read(), ["-F", "-lg", "40"])
    names = [f for f in os.listdir(GOLDEN) if f.startswith("demo_F_lg40__")]
    assert len(names) == 3
    for g in names:
        assert files[g.split("__", 1)[1]] == open(os.path.join(GOLDEN, g)).read(), g
    assert len(files["hg19-H19-testDNA-TFOsorted"].splitlines()) > 100


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_1mbp(engine):
    """Size-independent properties at bench scale (1 Mbp x 3 kb): determinism, shard-invariance (scanning the region as
    four shards cut at segment starts gives the same triplexes) and oracle spot checks on sampled tasks."""
    n = 1_000_000
    dna = splitmix_bases(1001, n)
    rna = splitmix_bases(2001, 3000)
    engine.set_params()
    engine.set_query("synRNA3k", rna)
    a = engine.LongTarget(dna, "chr1", 1)
    b = engine.LongTarget(dna, "chr1", 1)
    assert a == b
    stride = 4900
    cuts = [0, 51 * stride, 102 * stride, 153 * stride, n]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        end = min(n, hi + 100) if hi < n else n          # each shard carries the 100-bp overlap of its last segment
        rows = engine.LongTarget(dna[lo:end], "chr1", 1)
        for r in rows:
            r["starj"] += lo; r["endj"] += lo; r["genomestart"] += lo; r["genomeend"] += lo
        # a shard's trailing partial segment re-appears as the head of the next shard's first segment: drop rows that
        # come from the trailing overlap-only segment
        parts.append(rows)
    key = lambda r: (r["starj"], r["endj"], r["stari"], r["endi"], r["rule"], r["strand"], r["reverse"], f2b(r["score"]))
    whole = sorted(set(map(key, a)))
    shard = sorted(set(k for rows in parts for k in map(key, rows)))
    missing = [k for k in whole if k not in set(shard)]
    assert not missing, missing[:3]
    rnd = random.Random(3)
    for _ in range(6):
        s = rnd.randrange(0, n // stride) * stride
        seg = dna[s:s + 5000]
        tasks = rnd.sample(TASKS, 4)
        pr = engine.probe_segment(seg, tasks)
        for (pa, st, ru), r in zip(tasks, pr):
            s2, _ = O.task_strings(seg, pa, st, ru)
            mx = O.calc_score_once(rna, s2)
            assert r["max_score"] == mx and r["peaks"] == O.prealign(rna, s2, int(mx * 0.8))
            assert (r["colmax"] == O.colmax(rna, s2)).all()
    # three whole segments against the oracle, record level
    for s in (0, 49 * stride, 150 * stride):
        seg = dna[s:s + 5000]
        assert rows_as_oracle_text(engine.LongTarget(seg)) == oracle_text_rows(O.longtarget(rna, seg))


def test_shards_reproduce_the_whole_record(engine):
    """Multi-GPU path on one GPU: scanning a record as 3 shards through ltg_scan_shard (what each rank of bench.py --gpus N
    does) and concatenating the results in shard order gives exactly the rows of the unsharded scan."""
    n = 160_000
    dna = splitmix_bases(1001, n)
    rna = splitmix_bases(2001, 1200)
    engine.set_params(c_length=25)
    engine.set_query("synRNA", rna)
    whole = engine.LongTarget(dna, "chr1", 7)
    world, parts = 3, []
    for r in range(world):
        first, count, first_byte, n_bytes = fb.shard_segments(n, world, r)
        res = engine.scan_shard(dna[first_byte:first_byte + n_bytes], n, first, count, "chr1", 7)
        parts.append(fb.result_rows(res))
        engine.free(res)
    assert fb.merge_shard_rows(parts) == whole and len(whole) > 0


def test_row_pruning_is_exact(data_dir):
    """The window stage skips RNA rows that provably cannot hold a window's best cell (window.cuh).  With LTG_NO_PRUNE=1
    every window sweeps the whole lncRNA like the reference does: both modes must give identical rows, on random DNA
    (few strips survive) and on the repeat-rich demo (many ties, overflow and literal windows)."""
    dna = splitmix_bases(1001, 300_000)
    rna = splitmix_bases(2001, 3000)
    drna, _, ddna = demo(data_dir)
    os.environ["LTG_NO_PRUNE"] = "1"
    os.environ["LTG_NO_DEAD"] = "1"          # and trace every alignment, not only those that can still be reported
    os.environ["LTG_NO_SKIP"] = "1"          # and run every window round
    os.environ["LTG_NO_Q4PROBE"] = "1"       # and send every task that reaches 148 through the literal emulation
    try:
        full = fb.Engine(0)
    finally:
        del os.environ["LTG_NO_PRUNE"]
        del os.environ["LTG_NO_DEAD"]
        del os.environ["LTG_NO_SKIP"]
        del os.environ["LTG_NO_Q4PROBE"]
    pruned = fb.Engine(0)
    try:
        outs = []
        for eng in (full, pruned):
            eng.set_params(c_length=20)
            eng.set_query("synRNA3k", rna)
            res = eng.scan_record(dna, "chr1", 1)
            rows, cells = fb.result_rows(res), res.contents.window_cells
            eng.free(res)
            eng.set_query("H19", drna)
            rows2 = eng.LongTarget(ddna, "chr11", 1)
            # the dead-alignment rule under other filter settings: positive penaltyC, mild penaltyT, default cLength
            eng.set_params(penalty_c=1, penalty_t=-3, min_stability=2, min_identity=50)
            eng.set_query("synRNA3k", rna)
            rows3 = eng.LongTarget(dna[:150_000], "chr1", 1)
            eng.set_params(penalty_t=-500, penalty_c=1, nt_min=25, min_identity=70, c_length=60, c_distance=10)
            eng.set_query("H19", drna)
            rows4 = eng.LongTarget(ddna, "chr11", 1)
            outs.append((rows, rows2, cells, rows3, rows4))
        assert outs[0][0] == outs[1][0] and len(outs[0][0]) > 0
        assert outs[0][1] == outs[1][1] and len(outs[0][1]) > 0
        assert outs[0][3] == outs[1][3] and outs[0][4] == outs[1][4] and len(outs[0][4]) > 0
        assert outs[1][2] * 2 < outs[0][2]          # and it actually prunes: fewer than half of the window cells
    finally:
        full.close()
        pruned.close()


def test_q4_probe_is_exact(engine):
    """Tasks whose maximum reaches 148 are swept a second time to see whether an F >= 132 is carried into a stripe start of the
    reference's layout; only those go through the literal emulation.  With LTG_NO_Q4PROBE=1 all of them do: same rows, on a
    lncRNA with low-complexity tracts against repeat-rich DNA (most tasks reach 148), and far fewer literal tasks."""
    import random
    rng = random.Random(11)
    mut = lambda s, rate: "".join(rng.choice("ACGT") if rng.random() < rate else c for c in s)
    rna = list(splitmix_bases(2001, 2000))
    for at, unit in ((300, "CT"), (900, "GA"), (1500, "GT")):
        rna[at:at + 180] = mut(unit * 90, 0.08)
    rna = "".join(rna)
    dna = list(splitmix_bases(1001, 120_000))
    for _ in range(60):
        at = rng.randrange(0, len(dna) - 300)
        L = rng.randrange(30, 220)
        dna[at:at + L] = mut((rng.choice(["GA", "CT", "GAA", "CCT", "GGA", "TC"]) * L)[:L], rng.choice([0.0, 0.06, 0.12]))
    dna = "".join(dna)
    os.environ["LTG_NO_Q4PROBE"] = "1"
    try:
        plain = fb.Engine(0)
    finally:
        del os.environ["LTG_NO_Q4PROBE"]
    os.environ["LTG_LIT_OLD"] = "1"              # and the one-half-warp-per-job literal kernel instead of the column-parallel one
    try:
        old_kernel = fb.Engine(0)
    finally:
        del os.environ["LTG_LIT_OLD"]
    os.environ["LTG_FREC"] = "1"                 # the main sweep records a stripe-start screen (lane resolution) next to the granule pre-filter
    try:
        recording = fb.Engine(0)
    finally:
        del os.environ["LTG_FREC"]
    os.environ["LTG_Q4_TAINT"] = "0"             # round 1's path: probe sweep (largest carried F) instead of the taint sweep
    os.environ["LTG_FREC"] = "0"
    try:
        probing = fb.Engine(0)
    finally:
        del os.environ["LTG_Q4_TAINT"]
        del os.environ["LTG_FREC"]
    os.environ["LTG_Q4_TAINT"] = "0"             # ... and the screen alone (no second sweep at all)
    os.environ["LTG_FREC"] = "1"
    try:
        recording_only = fb.Engine(0)
    finally:
        del os.environ["LTG_Q4_TAINT"]
        del os.environ["LTG_FREC"]
    os.environ["LTG_WIN_Q4CHK"] = "1"            # window sweeps watch the stripe starts: only the windows that saw an F >= 132 there are emulated
    try:
        win_checked = fb.Engine(0)
    finally:
        del os.environ["LTG_WIN_Q4CHK"]
    os.environ["LTG_WIN_Q4CHK"] = "0"
    try:
        win_unchecked = fb.Engine(0)
    finally:
        del os.environ["LTG_WIN_Q4CHK"]
    engines = (plain, engine, old_kernel, recording, probing, recording_only, win_checked, win_unchecked)
    try:
        outs, lit_windows = [], []
        for eng in engines:
            eng.set_params(c_length=25)
            eng.set_query("lnc", rna)
            res = eng.scan_record(dna, "chr1", 1)
            outs.append((fb.result_rows(res), res.contents.n_literal_tasks, res.contents.n_q4_probed))
            lit_windows.append(res.contents.n_literal_windows)
            eng.free(res)
        assert 0 < lit_windows[6] < lit_windows[7] // 2 and lit_windows[7] > 100
        rows = outs[0][0]
        assert len(rows) > 100
        for o in outs[1:]:
            assert o[0] == rows
        assert outs[0][2] == 0 and outs[1][2] > 0
        assert outs[2] == outs[1]
        # literal tasks: all flagged (plain) > probe sweep >= taint sweep (it certifies most of what the probe leaves)
        assert 0 < outs[4][1] < outs[0][1] and outs[1][1] < outs[4][1]
        # stripe-start screen recorded by the main sweep, alone: no second sweep; coarser than the probe, finer than no filter
        assert outs[5][2] == 0 and outs[4][1] <= outs[5][1] < outs[0][1]
        # screen + taint: fewer pairs reach the taint sweep than with the granule pre-filter alone, and no more literal tasks
        assert 0 < outs[3][2] < outs[1][2] and outs[3][1] <= outs[1][1]
    finally:
        for eng in engines:
            if eng is not engine:
                eng.close()
        engine.set_params()


def test_window_q4_check_on_device():
    """Window sweeps that watch the stripe starts (k_win_dp Q4CHK, forced on): only windows that saw an F >= 132 enter a stripe
    start of their alignment call are emulated literally.  Planted windows (insertions that start at a stripe start, the
    constellation that makes the Q4 quirk visible in Aligner::Align; tests/test_q4_theory_cpu.py) through the function-level
    seam: score and the four coordinates must equal the reference's in every case, including the ones where the reference
    really deviates from exact Smith-Waterman — those must have been flagged — and the constructed reproducer.  Runs in a
    process of its own (tests/_window_q4_check_job.py): ltg_probe_align needs the device's task tables for itself."""
    env = dict(os.environ, LTG_WIN_Q4CHK="1")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_window_q4_check_job.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_q4_taint_certification_on_device(engine):
    """The taint sweep (k_scan TAINT) returns flagged tasks to the exact path only when their recorded column maxima are
    provably the reference's.  Planted cases (insertions that start exactly at a stripe start, the constellation that makes the
    Q4 quirk visible; tests/test_q4_theory_cpu.py) through the function-level seam: whatever the verdict, the column maxima must
    equal the reference's; most flagged cases must be certified; every really different case must stay literal.  The verdicts are
    also compared with the CPU prototype of the same rule (they may differ on taint ties, which both resolve soundly)."""
    import random
    from test_q4_theory_cpu import certify, exact_colmax_and_carried_f, make_case
    S = ref_side() if have_ref_shim() else oracle_side()
    inv = str.maketrans("GCTA", "ACGT")           # AntiMinus rule 8 maps A->G C->C G->T T->A (a bijection): raw DNA of a translated string
    task = (-1, 1, 8)
    rng = random.Random(5)
    cases = [make_case(rng) for _ in range(160)]
    # longer lncRNAs: several strips of the scan kernel (R = 32: 1024 rows per strip), stripe starts inside any lane
    for _ in range(40):
        m = rng.randrange(1100, 3300)
        n = rng.randrange(100, 400)
        L = (m + 15) // 16
        rna = [rng.choice("ACGT") for _ in range(m)]
        dna = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(rng.randrange(1, 4)):
            b = L * rng.randrange(1, 16)
            a = max(0, b - rng.randrange(28, 48))
            gap = rng.randrange(2, 7)
            frag = rna[a:b] + rna[b + gap:b + gap + rng.randrange(8, 30)]
            frag = [c if rng.random() > 0.03 else rng.choice("ACGT") for c in frag]
            at = rng.randrange(0, max(1, n - len(frag)))
            dna[at:at + len(frag)] = frag
        cases.append(("".join(rna), "".join(dna[:n])))
    engine.set_params()
    flagged = certified = different = agree = 0
    for rna, dna_t in cases:
        raw = dna_t.translate(inv)
        assert S.task_strings(raw, *task)[0] == dna_t
        engine.set_query("lnc", rna)
        got = engine.probe_segment(raw, [task])[0]
        ref = S.colmax(rna, dna_t)
        assert (got["colmax"] == ref).all(), (rna, dna_t)
        exact, fmax = exact_colmax_and_carried_f(rna, dna_t)
        differs = not np.array_equal(exact, ref)
        if fmax >= 132:
            flagged += 1
            certified += 0 if got["literal"] else 1
            different += int(differs)
            assert not (differs and not got["literal"]), (rna, dna_t)
            agree += int(bool(got["literal"]) == (not certify(rna, dna_t, kernel_rule=True)))
        else:
            assert not differs
    print("flagged %d, certified on the device %d, really different %d, same verdict as the CPU prototype %d" % (flagged, certified, different, agree))
    assert flagged > 120 and different >= 5
    assert certified >= flagged // 2
    assert agree >= flagged - flagged // 10


def test_error_behaviour(engine):
    with pytest.raises(fb.FasimError):
        engine.set_params(rule=19, strand=-1)           # reference: exit(1) in transferString
    with pytest.raises(fb.FasimError):
        engine.set_params(cut_length=100, overlap=100)  # reference: cutSequence never advances
    engine.set_params()
