#!/usr/bin/env python
"""Golden files of the `--compat lowercase` mode: the UNMODIFIED older variant (fasim-LongTarget.cpp + fastSim.h, built by
oracle/Makefile as oracle/_ref/fasim_lc) on the demo (-lg 40) and on the first 12 MEG3 example regions (-lg 60, one multi-record
file — that variant's reader handles several records).  Run in the build container only."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from _harness import REF_DIR  # noqa: E402
from make_golden import DATA, run_ref  # noqa: E402


def main():
    binary = os.path.join(REF_DIR, "fasim_lc")
    rd = lambda n: open(os.path.join(DATA, n)).read()
    files = run_ref(binary, "testDNA.fa", rd("testDNA.fa"), "H19.fa", rd("H19.fa"), ["-lg", "40"])
    assert list(files) == ["hg19-H19-fastSim-TFOsorted"], sorted(files)
    open(os.path.join(HERE, "demo_lc_lg40__hg19-H19-fastSim-TFOsorted"), "w").write(files["hg19-H19-fastSim-TFOsorted"])
    files = run_ref(binary, "MEG3-12.fa", rd("MEG3-DNAseq-first12.fa"), "MEG3.fa", rd("MEG3-ENST00000451743.fa"), ["-lg", "60"])
    (name, text), = files.items()
    open(os.path.join(HERE, "meg3_first12_lc__" + name), "w").write(text)
    print(name, len(text.splitlines()), "lines")


if __name__ == "__main__":
    main()
