#!/bin/bash
# Round 2, visit H: remaining new tests, the full parity suite, ncu captures (tools/gpu_r2f_ncu.sh), headline bench.
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "sim_mode_vs_oracle or packed or envelope" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?" >> gpurun_out/${TAG}_pytest_new.log
tail -n 8 gpurun_out/${TAG}_pytest_new.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
bash tools/gpu_r2f_ncu.sh ${TAG}
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench.json'))[-1:]:
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f e2e=%.0f ms=%.1f frac=%.3f parity=%s live=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'], j.get('parity_sample',{}).get('equal'), str(j['roofline'].get('live_ubench'))[:300]))
P
