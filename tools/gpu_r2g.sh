#!/bin/bash
# Round 2, visit G: -F kernel with warp-parallel phase B — tests and wall time; packed / envelope tests.
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sim_mode_vs_oracle or packed or envelope" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?" >> gpurun_out/${TAG}_pytest_new.log
tail -n 25 gpurun_out/${TAG}_pytest_new.log | cut -c1-300
( cd tests/golden/data && mkdir -p /tmp/fo && time ( LTG_TIMING=1 timeout 600 ../../../fasim-longtarget_b200/fasim -f1 testDNA.fa -f2 H19.fa -O /tmp/fo/ -F -lg 40 ) ) > gpurun_out/${TAG}_demoF.log 2>&1; grep -E "real|Running time|finished|timing" gpurun_out/${TAG}_demoF.log
cmp /tmp/fo/hg19-H19-testDNA-TFOsorted tests/golden/demo_F_lg40__hg19-H19-testDNA-TFOsorted && echo "demo -F byte-equal"
