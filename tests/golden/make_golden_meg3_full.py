#!/usr/bin/env python
"""Golden for BASELINE.json configs[1] — the MEG3 lncRNA against all 532 example regions with default flags
(-c 5000 -i 60 are the defaults) — from the reference itself.

The canonical reader cannot parse multi-record FASTA, so the whole-file run uses oracle/_ref/fasim_mr:
the unmodified sources plus the one-line accumulator reset in readDna (oracle/Makefile, "reference + multi-record
fix").  Build container only (needs /root/reference); ~2 CPU-minutes.  Outputs are stored gzip-compressed.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _harness import REF_DIR  # noqa: E402

REF = "/root/reference/example/MEG3"


def main():
    d = tempfile.mkdtemp()
    try:
        shutil.copyfile(os.path.join(REF, "MEG3-ENST00000451743-DNAseq.fa"), os.path.join(d, "MEG3-DNAseq.fa"))
        shutil.copyfile(os.path.join(REF, "MEG3-ENST00000451743.fa"), os.path.join(d, "MEG3.fa"))
        os.mkdir(os.path.join(d, "out"))
        subprocess.run([os.path.join(REF_DIR, "fasim_mr"), "-f1", "MEG3-DNAseq.fa", "-f2", "MEG3.fa", "-O", "out/"], cwd=d,
                       stdout=subprocess.DEVNULL, check=True, timeout=3600)
        outs = sorted(os.listdir(os.path.join(d, "out")))
        sorted_name = [f for f in outs if f.endswith("TFOsorted")][0]
        with open(os.path.join(d, "MEG3-DNAseq.fa"), "rb") as src, gzip.GzipFile(os.path.join(HERE, "data", "MEG3-DNAseq.fa.gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        with open(os.path.join(d, "out", sorted_name), "rb") as src, gzip.GzipFile(os.path.join(HERE, "meg3_full_mr_defaults__TFOsorted.gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        print(sorted_name, sum(1 for _ in open(os.path.join(d, "out", sorted_name))), "lines")
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
