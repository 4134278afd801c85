#!/bin/bash
# Round 2, final visit: headline bench with the reference beside it, reference arm, real-data and multi-query lines, launch list,
# ncu --set full captures (plain scan, taint sweep, window sweep), smoke.  Bench values are never taken under ncu.
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
for cfg in demo meg3 h19 malat1 neat1; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${cfg}.json 2> gpurun_out/${TAG}_bench_${cfg}.err; echo "$cfg rc=$?"
done
timeout 600 python bench.py --queries 8 --region-mbp 10 --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_mq8.json 2> gpurun_out/${TAG}_bench_mq8.err; echo "mq8 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${TAG}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
B="python bench.py --steps 1 --warmup 1 --region-mbp 2.5 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_scan -c 1 -f -o gpurun_out/${TAG}_scan $B > gpurun_out/${TAG}_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_win_dp -c 1 -f -o gpurun_out/${TAG}_windp $B > gpurun_out/${TAG}_ncu_windp.log 2>&1; echo "ncu windp rc=$?"
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:k_scan.*0, 1>' -c 1 -f -o gpurun_out/${TAG}_taint python bench.py --config h19 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_taint.log 2>&1; echo "ncu taint rc=$?"
for n in scan windp taint; do
  ncu -i gpurun_out/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${TAG}_${n}_raw.csv > gpurun_out/${TAG}_${n}_summary.md 2>&1
done
rm -f gpurun_out/${TAG}_*.ncu-rep
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_bench*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.0f e2e=%.0f ms=%.1f frac=%s win=%s parity=%s cpu=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j.get('roofline',{}).get('frac'), j.get('stage_ms_per_step',{}).get('window'), j.get('parity_sample',{}).get('equal'), j.get('cpu_baseline',{}).get('value')))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do echo "$f: $(tail -n 1 $f | cut -c1-300)"; done
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -30
head -30 gpurun_out/${TAG}_taint_summary.md
