#!/bin/bash
# Round 2, visit R: validation of the final tree — parity suite as shipped, parity suite with every optional Q4 variant forced on
# (window stripe-start check, stripe-start screen in the main sweep), headline bench, smoke.
TAG=${1:-r02r}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 8 gpurun_out/${TAG}_pytest.log | cut -c1-400
LTG_WIN_Q4CHK=1 LTG_FREC=1 timeout 2400 python -m pytest tests -m gpu -q -k "not q4_probe_is_exact" > gpurun_out/${TAG}_pytest_forced.log 2>&1; echo "pytest forced rc=$?" >> gpurun_out/${TAG}_pytest_forced.log
tail -n 8 gpurun_out/${TAG}_pytest_forced.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${TAG}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --region-mbp 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02r_bench*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.0f e2e=%.0f ms=%.1f frac=%s win=%s parity=%s cpu=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j.get('roofline',{}).get('frac'), j.get('stage_ms_per_step',{}).get('window'), j.get('parity_sample',{}).get('equal'), j.get('cpu_baseline',{}).get('value')))
    except Exception as e: print(f,'ERR',e)
P
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | grep -E "launches|k_scan|k_win|k_epi"
