#!/bin/bash
# Round 2, visit T: the final tree — full parity suite (the window-check test now runs in a process of its own), smoke.
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 8 gpurun_out/${TAG}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
