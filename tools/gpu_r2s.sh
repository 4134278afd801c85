#!/bin/bash
# Round 2, visit S: the window stripe-start check through the function-level seam (planted windows incl. real deviations).
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "window_q4_check" > gpurun_out/${TAG}_pytest_wincheck.log 2>&1; echo "pytest wincheck rc=$?"
tail -n 12 gpurun_out/${TAG}_pytest_wincheck.log | cut -c1-400
LTG_WIN_Q4CHK=1 timeout 900 python -m pytest tests -m gpu -q -k "window_alignments_demo or q4_reproducer or microsatellite or tiny_and_ragged or q4_probe_is_exact" > gpurun_out/${TAG}_pytest_forced.log 2>&1; echo "pytest forced rc=$?"
tail -n 8 gpurun_out/${TAG}_pytest_forced.log | cut -c1-400
