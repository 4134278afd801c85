#!/bin/bash
# Round 2, ncu visit: full captures of k_scan (shipped and split-E variant), k_win_dp forward; raw CSV pages exported on the box
# (the .ncu-rep files come back too).  Bench values are never taken under ncu.
TAG=${1:-r02f}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --region-mbp 2.5 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_scan -c 1 -f -o gpurun_out/${TAG}_scan $B > gpurun_out/${TAG}_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
FASIM_B200_LIB=fasim-longtarget_b200/variants/libfasim_b200_splitE.so ncu --set full --clock-control none -k regex:k_scan -c 1 -f -o gpurun_out/${TAG}_scan_splitE $B > gpurun_out/${TAG}_ncu_scan_splitE.log 2>&1; echo "ncu splitE rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_win_dp -c 1 -f -o gpurun_out/${TAG}_windp $B > gpurun_out/${TAG}_ncu_windp.log 2>&1; echo "ncu windp rc=$?"
for n in scan scan_splitE windp; do
  ncu -i gpurun_out/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${TAG}_${n}_raw.csv > gpurun_out/${TAG}_${n}_summary.md 2>&1
done
cat gpurun_out/${TAG}_scan_summary.md | head -30
# split-E variant timed without ncu (scan GCUPS of the two builds side by side)
timeout 300 python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline > gpurun_out/${TAG}_ab_default.json 2> /dev/null
FASIM_B200_LIB=fasim-longtarget_b200/variants/libfasim_b200_splitE.so timeout 300 python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline > gpurun_out/${TAG}_ab_splitE.json 2> /dev/null
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_ab_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f scan_gcups=%.0f'%(j['value'], j['roofline']['achieved']))
    except Exception as e: print(f,'ERR',e)
P
