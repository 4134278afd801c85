#!/usr/bin/env python
"""Golden files of the -F (SIM) mode: the UNMODIFIED reference binary run with -F on the H19 / testDNA demo (-lg 40).
Run in the build container only (needs oracle/_ref/fasim, i.e. /root/reference); takes ~2 minutes of CPU."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from _harness import ref_binary  # noqa: E402
from make_golden import DATA, run_ref  # noqa: E402


def main():
    files = run_ref(ref_binary(), "testDNA.fa", open(os.path.join(DATA, "testDNA.fa")).read(), "H19.fa",
                    open(os.path.join(DATA, "H19.fa")).read(), ["-F", "-lg", "40"])
    assert len(files) == 3, sorted(files)
    for name, text in files.items():
        open(os.path.join(HERE, "demo_F_lg40__" + name), "w").write(text)
        print(name, len(text.splitlines()), "lines")


if __name__ == "__main__":
    main()
