#!/bin/bash
# Round 2, visit D: -F (SIM) and --compat lowercase tests, the full parity suite, wall time of `fasim -F` on the demo.
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sim_mode or compat" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?" >> gpurun_out/${TAG}_pytest_new.log
tail -n 30 gpurun_out/${TAG}_pytest_new.log | cut -c1-300
( cd tests/golden/data && mkdir -p /tmp/fo && time ( LTG_TIMING=1 ../../../fasim-longtarget_b200/fasim -f1 testDNA.fa -f2 H19.fa -O /tmp/fo/ -F -lg 40 ) ) > gpurun_out/${TAG}_demoF.log 2>&1; grep -E "real|Running time|finished|timing" gpurun_out/${TAG}_demoF.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
