#!/bin/bash
# Round 2, visit A: parity tests (incl. the new reference-binary tests), headline bench with parity_sample, real-data lines.
TAG=${1:-r02a}
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv; } > gpurun_out/${TAG}_box.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${TAG}_pytest.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
for cfg in demo meg3 h19 malat1 neat1; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_$cfg.json 2> gpurun_out/${TAG}_bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 900 python bench.py --queries 8 --region-mbp 10 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_mq8.json 2> gpurun_out/${TAG}_bench_mq8.err; echo "bench mq8 rc=$?"
cat gpurun_out/${TAG}_bench*.json | cut -c1-1500
for f in gpurun_out/${TAG}_bench*.err; do tail -n 3 $f; done
