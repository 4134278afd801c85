#!/bin/bash
# Round-end style measurement on the GPU box: parity tests, the default bench (100 Mbp), the reference arm, the launch
# list of a short run and one full ncu capture of the dominant kernels.  Usage: gpurun -- bash tools/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan -c 1 -f -o gpurun_out/${TAG}_scan python bench.py --steps 1 --warmup 1 --region-mbp 2.5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_win_dp -c 1 -f -o gpurun_out/${TAG}_windp python bench.py --steps 1 --warmup 1 --region-mbp 2.5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_windp.log 2>&1
cat gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_reference.json
