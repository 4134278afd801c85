"""CPU: the C-ABI library builds, loads and exports every symbol include/fasim_b200.h declares; without a GPU the
product fails loudly (no CPU fallback); host-side pieces of the ABI (cluster / writer) match the oracle."""
import ctypes as C
import os
import re

import pytest

import fasim_b200 as fb
from _harness import ROOT, oracle_side


def test_library_builds_and_exports_header_symbols():
    fb.build()
    hdr = open(os.path.join(ROOT, "include", "fasim_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ltg_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 18
    L = fb.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(fb.EXPORTS) == declared


def test_struct_layouts_match_header():
    assert C.sizeof(fb.Params) == 48
    assert C.sizeof(fb.Triplex) == 104
    assert C.sizeof(fb.TaskProbe) == 32


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(fb.FasimError) as e:
        fb.Engine(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_oracle_is_not_linked_into_the_product():
    import subprocess
    out = subprocess.run(["ldd", fb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ref_shim" not in out
    src_dir = os.path.join(ROOT, "fasim-longtarget_b200")
    for dp, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".inl", ".py", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "oracle/" not in txt and "orc_" not in txt, f


def _make_result(rows, chr_tag="chrT"):
    """Hand-made ltg_result (host memory) to exercise ltg_cluster / ltg_write_tfosorted without a GPU."""
    L = fb.lib()
    n = len(rows)
    text = bytearray(chr_tag.encode() + b"\0")
    arr = (fb.Triplex * max(n, 1))()
    for i, r in enumerate(rows):
        t = arr[i]
        for k in ("stari", "endi", "starj", "endj", "reverse", "strand", "rule", "nt"):
            setattr(t, k, r[k])
        t.score, t.identity, t.tri_score = r["score"], r["identity"], r["tri_score"]
        t.genomestart, t.genomeend = r["starj"] + 999, r["endj"] + 999
        t.tfo_off = len(text); text += r["tfo"].encode() + b"\0"
        t.tts_off = len(text); text += r["tts"].encode() + b"\0"
        t.chr_off = 0
    src = fb.Result()
    src.n_triplex = n
    src.triplex = C.cast(arr, C.POINTER(fb.Triplex))
    buf = C.create_string_buffer(bytes(text), len(text))
    src.text_bytes = len(text)
    src.text = C.cast(buf, C.POINTER(C.c_char))
    dst = C.POINTER(fb.Result)()
    assert L.ltg_result_new(C.byref(dst)) == 0
    assert L.ltg_result_append(dst, C.byref(src)) == 0
    return dst, (arr, buf)


def test_cluster_and_writer_match_oracle(tmp_path):
    import random
    rnd = random.Random(5)
    rows = []
    for i in range(60):
        a = rnd.randrange(20, 900)
        ln = rnd.randrange(30, 90)
        b = rnd.randrange(1000, 4000)
        fwd = rnd.random() < 0.5
        rows.append(dict(stari=a, endi=a + ln, starj=b if fwd else b + ln, endj=b + ln if fwd else b, reverse=rnd.choice([1, -1]),
                         strand=rnd.choice([0, 1]), rule=rnd.randrange(1, 7), nt=ln + rnd.randrange(0, 3), score=float(rnd.randrange(60, 200)),
                         identity=rnd.uniform(60, 100), tri_score=rnd.uniform(1, 3), tfo="ACGT-" * 3, tts="TG-CA" * 3))
    res, keep = _make_result(rows)
    p = fb.default_params(c_distance=15, c_length=50)
    assert fb.lib().ltg_cluster(res, C.byref(p)) == 0
    got = fb.result_rows(res)
    mid, cen, mot = oracle_side().cluster([r["stari"] for r in rows], [r["endi"] for r in rows], [r["nt"] for r in rows], 15, 50)
    # the product result is sorted by class (unstable std::sort): compare as a multiset of (identity key -> class triple)
    exp = sorted((r["stari"], r["endi"], r["starj"], r["nt"], m, c, k) for r, m, c, k in zip(rows, mid, cen, mot))
    have = sorted((r["stari"], r["endi"], r["starj"], r["nt"], r["middle"], r["center"], r["motif"]) for r in got)
    assert exp == have
    assert [r["motif"] for r in got] == sorted(r["motif"] for r in got)
    out = str(tmp_path / "x-TFOsorted")
    assert fb.lib().ltg_write_tfosorted(res, out.encode()) == 0
    lines = open(out).read().splitlines()
    assert lines[0].split("\t")[:4] == ["QueryStart", "QueryEnd", "StartInSeq", "EndInSeq"] and len(lines[0].split("\t")) == 19
    body = [l.split("\t") for l in lines[1:]]
    assert len(body) == sum(1 for r in got if r["motif"] != 0)
    for f, r in zip(body, [r for r in got if r["motif"] != 0]):
        assert f[4] == ("R" if r["starj"] < r["endj"] else "L") and f[5] == "chrT"
        assert f[8] == "%g" % r["tri_score"] and f[9] == "%g" % r["identity"] and f[12] == "%g" % r["score"]
    fb.lib().ltg_result_free(res)


def test_cli_usage_and_bad_flags():
    fb.build()
    r = fb.run_cli([])
    assert r.returncode == 1 and "-f1" in r.stdout
