"""Body of tests/test_gpu_parity.py::test_window_q4_check_on_device (run as a script with LTG_WIN_Q4CHK=1 in a process of its own)."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "fasim-longtarget_b200"))

import fasim_b200 as fb
from _harness import GOLDEN, have_ref_shim, oracle_side, ref_side
from test_q4_theory_cpu import exact_align, make_case


def main():
    assert os.environ.get("LTG_WIN_Q4CHK") == "1"
    S = ref_side() if have_ref_shim() else oracle_side()
    golden = json.load(open(os.path.join(GOLDEN, "golden.json")))
    eng = fb.Engine(0)
    try:
        eng.set_params()
        q = golden["q4"]
        eng.set_query("q4", q["rna"])
        (o5, cig), = eng.Align([q["dna"]])
        assert [list(o5), cig] == [q["align"][0], [c for c in q["align"][1] if c >> 4]] and o5[0] == 181
        rng = random.Random(77)
        n_cases = n_high = n_diverged = 0
        for _ in range(500):
            rna, win = make_case(rng)
            win = win[:196]
            ex = exact_align(rna, win)
            if ex is None:
                continue
            want_exact, fmax = ex
            ref5, _ = S.align(rna, win)
            eng.set_query("lnc", rna)
            (got5, _), = eng.Align([win])
            assert tuple(got5) == tuple(ref5), (rna, win, got5, ref5, fmax)
            n_cases += 1
            n_high += int(want_exact[0] >= 148)
            n_diverged += int(tuple(ref5) != tuple(want_exact))
        print("windows %d, reach 148: %d, reference differs from exact SW: %d" % (n_cases, n_high, n_diverged))
        assert n_cases > 300 and n_high > 100 and n_diverged >= 5
    finally:
        eng.close()


if __name__ == "__main__":
    main()
