#!/usr/bin/env python
"""Golden for BASELINE.json configs[2] on the second substitute DNA set (the NEAT1 / MALAT1 example DNA
files are absent from the reference checkout, so the lncRNAs are run against testDNA.fa — make_golden.py — and against the
MEG3 example regions — this script, first 12 regions) with the complex flags.  Reference = oracle/_ref/fasim_mr (unmodified
sources + the one-line multi-record reader fix, oracle/Makefile).  Build container only."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _harness import REF_DIR  # noqa: E402

COMPLEX2 = ["-i", "70", "-S", "1.0", "-ni", "25", "-pt", "-500", "-ds", "10", "-lg", "60"]


def main():
    for name in ("NEAT1", "MALAT1"):
        d = tempfile.mkdtemp()
        try:
            shutil.copyfile(os.path.join(HERE, "data", "MEG3-DNAseq-first12.fa"), os.path.join(d, "MEG3-12.fa"))
            shutil.copyfile(os.path.join(HERE, "data", name + ".fa"), os.path.join(d, name + ".fa"))
            os.mkdir(os.path.join(d, "out"))
            subprocess.run([os.path.join(REF_DIR, "fasim_mr"), "-f1", "MEG3-12.fa", "-f2", name + ".fa", "-O", "out/"] + COMPLEX2, cwd=d,
                           stdout=subprocess.DEVNULL, check=True, timeout=3600)
            out = [f for f in os.listdir(os.path.join(d, "out")) if f.endswith("TFOsorted")][0]
            shutil.copyfile(os.path.join(d, "out", out), os.path.join(HERE, "%s_meg3first12_complex__TFOsorted" % name))
            print(name, out, sum(1 for _ in open(os.path.join(d, "out", out))), "lines")
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
