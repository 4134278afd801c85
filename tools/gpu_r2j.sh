#!/bin/bash
# Round 2, visit J: carried-F recording in the main sweep (FREC) + bulk-copy profile staging — parity suite, A/B against the
# previous build (variants/libfasim_b200_base.so), real-data configs with the probe (LTG_FREC=0) and adaptive, launch lists.
TAG=${1:-r02j}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
timeout 600 $B --debug-stats > gpurun_out/${TAG}_ab_new.json 2> gpurun_out/${TAG}_ab_new.err; echo "new rc=$?"
FASIM_B200_LIB=fasim-longtarget_b200/variants/libfasim_b200_base.so timeout 600 $B > gpurun_out/${TAG}_ab_base.json 2> gpurun_out/${TAG}_ab_base.err; echo "base rc=$?"
LTG_FREC=1 timeout 600 $B > gpurun_out/${TAG}_ab_frec1.json 2> gpurun_out/${TAG}_ab_frec1.err; echo "frec1 rc=$?"
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 8 gpurun_out/${TAG}_pytest.log | cut -c1-300
for cfg in neat1 h19 malat1; do
  timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_auto.json 2> gpurun_out/${TAG}_${cfg}_auto.err; echo "$cfg auto rc=$?"
  LTG_FREC=0 timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_probe.json 2> gpurun_out/${TAG}_${cfg}_probe.err; echo "$cfg probe rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_neat1.csv python bench.py --config neat1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_neat1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f e2e=%.0f ms=%.1f scan_gcups=%.0f scan_ms=%.0f win=%.0f lit=%s probed=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['achieved'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['literal_tasks_per_step'], j['q4_probed_pairs_per_step']))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do tail -n 2 $f | cut -c1-400; done
python tools/launch_summary.py gpurun_out/${TAG}_launches_neat1.csv | head -30
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -30
