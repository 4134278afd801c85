#!/bin/bash
# Round 2, visit C: -F (SIM) kernel tests, full parity suite, headline bench after the k_win_dp / planner changes, launch list.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sim_mode" > gpurun_out/${TAG}_pytest_sim.log 2>&1; echo "pytest sim rc=$?" >> gpurun_out/${TAG}_pytest_sim.log
tail -n 25 gpurun_out/${TAG}_pytest_sim.log | cut -c1-400
( cd tests/golden/data && mkdir -p /tmp/fo && /usr/bin/time -v ../../../fasim-longtarget_b200/fasim -f1 testDNA.fa -f2 H19.fa -O /tmp/fo/ -F -lg 40 ) > gpurun_out/${TAG}_demoF.log 2>&1; grep -E "Elapsed|Running time|finished" gpurun_out/${TAG}_demoF.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --debug-stats > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
LTG_FLOORS=1 timeout 600 $B > gpurun_out/${TAG}_ab_floors.json 2> gpurun_out/${TAG}_ab_floors.err; echo "floors rc=$?"
timeout 600 $B > gpurun_out/${TAG}_ab_default.json 2> gpurun_out/${TAG}_ab_default.err; echo "default rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
for f in gpurun_out/${TAG}_*.json; do echo $f; python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value=%.0f e2e=%.0f ms=%.1f scan=%.0f win=%.0f frac=%.3f parity=%s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['roofline']['frac'], j.get('parity_sample',{}).get('equal')))
except Exception as e: print(' ERR', e)
P
done
for f in gpurun_out/${TAG}_*.err; do echo $f; tail -n 2 $f | cut -c1-600; done
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -30
