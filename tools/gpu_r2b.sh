#!/bin/bash
# Round 2, visit B: block maxima + accept floor.  Parity tests, A/B of the tracker floor and batch sizes, launch list.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline --debug-stats"
timeout 600 $B > gpurun_out/${TAG}_ab_default.json 2> gpurun_out/${TAG}_ab_default.err; echo "default rc=$?"
LTG_FLOORS=1 timeout 600 $B > gpurun_out/${TAG}_ab_floors.json 2> gpurun_out/${TAG}_ab_floors.err; echo "floors rc=$?"
LTG_BATCH_SEGMENTS=1024 timeout 600 $B > gpurun_out/${TAG}_ab_bs1024.json 2> gpurun_out/${TAG}_ab_bs1024.err; echo "bs1024 rc=$?"
LTG_BATCH_SEGMENTS=4096 timeout 600 $B > gpurun_out/${TAG}_ab_bs4096.json 2> gpurun_out/${TAG}_ab_bs4096.err; echo "bs4096 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --debug-stats > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
for f in gpurun_out/${TAG}_*.json; do echo $f; python - "$f" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value=%.0f e2e=%.0f ms=%.1f scan=%.0f win=%.0f frac=%.3f parity=%s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['roofline']['frac'], j.get('parity_sample',{}).get('equal')))
P
done
for f in gpurun_out/${TAG}_*.err; do echo $f; tail -n 2 $f | cut -c1-1200; done
