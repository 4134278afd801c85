#!/bin/bash
# Round 2, visit Q: the recording sweep as a branch-free stripe-start screen (lane resolution), switched on per query.
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 14 gpurun_out/${TAG}_pytest.log | cut -c1-400
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
timeout 600 $B > gpurun_out/${TAG}_ab_new.json 2> gpurun_out/${TAG}_ab_new.err; echo "new rc=$?"
LTG_FREC=1 timeout 600 $B > gpurun_out/${TAG}_ab_frec1.json 2> gpurun_out/${TAG}_ab_frec1.err; echo "frec1 rc=$?"
for cfg in neat1 h19 malat1 meg3; do
  timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_auto.json 2> gpurun_out/${TAG}_${cfg}_auto.err; echo "$cfg auto rc=$?"
  LTG_FREC=0 timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_nofrec.json 2> gpurun_out/${TAG}_${cfg}_nofrec.err; echo "$cfg nofrec rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_neat1.csv python bench.py --config neat1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_neat1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_h19.csv python bench.py --config h19 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_h19.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:k_scan.*\(bool\)1>' -c 1 -f -o gpurun_out/${TAG}_taint python bench.py --config h19 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_taint.log 2>&1; echo "ncu taint rc=$?"
ncu -i gpurun_out/${TAG}_taint.ncu-rep --page raw --csv > gpurun_out/${TAG}_taint_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/${TAG}_taint_raw.csv > gpurun_out/${TAG}_taint_summary.md 2>&1; rm -f gpurun_out/${TAG}_taint.ncu-rep
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02q_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f ms=%.1f scan_gcups=%.0f scan_ms=%.0f win=%.0f rows=%s lit=%s probed=%s'%(j['value'], j['ms_per_step'], j['roofline']['achieved'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['triplex_rows_per_step'], j['literal_tasks_per_step'], j['q4_probed_pairs_per_step']))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do echo "$f: $(tail -n 1 $f | cut -c1-300)"; done
python tools/launch_summary.py gpurun_out/${TAG}_launches_neat1.csv | grep -E "launches|k_scan|k_win_dp|k_literal|k_trace"
python tools/launch_summary.py gpurun_out/${TAG}_launches_h19.csv | grep -E "launches|k_scan|k_win_dp|k_literal|k_trace"
head -28 gpurun_out/${TAG}_taint_summary.md
