#!/bin/bash
# one GPU-box visit: parity tests, a short bench, host facts.  Usage: gpurun -- bash tools/gpu_round.sh [region_mbp]
REGION=${1:-10}
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/pytest_gpu.log 2>&1; echo "smoke rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 --region-mbp $REGION > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
cat gpurun_out/box.txt
if [ -n "$2" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/$2_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 5 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
fi
