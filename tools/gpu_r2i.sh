#!/bin/bash
# Round 2, visit I: shared-profile scan (6 warps per CTA, one profile copy) — parity suite, A/B against the per-warp variant, bench.
TAG=${1:-r02i}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
timeout 600 $B > gpurun_out/${TAG}_ab_shared.json 2> gpurun_out/${TAG}_ab_shared.err; echo "shared rc=$?"
LTG_SCAN_SHARED=0 timeout 600 $B > gpurun_out/${TAG}_ab_perwarp.json 2> gpurun_out/${TAG}_ab_perwarp.err; echo "perwarp rc=$?"
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --queries 8 --region-mbp 10 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_mq8.json 2> gpurun_out/${TAG}_bench_mq8.err; echo "bench mq8 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f e2e=%.0f ms=%.1f scan_gcups=%.0f win=%.0f parity=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['achieved'], j['stage_ms_per_step']['window'], j.get('parity_sample',{}).get('equal')))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do tail -n 2 $f | cut -c1-300; done
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -12
