#!/bin/bash
# Round 2, visit V: taint sweep ends at the last column a flagged task still records — full parity suite, real-data configs.
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 8 gpurun_out/${TAG}_pytest.log | cut -c1-300
for cfg in neat1 h19 malat1; do
  timeout 200 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}.json 2> gpurun_out/${TAG}_${cfg}.err; echo "$cfg rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02v_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f ms=%.1f scan_ms=%.0f win=%.0f rows=%s lit=%s probed=%s'%(j['value'], j['ms_per_step'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['triplex_rows_per_step'], j['literal_tasks_per_step'], j['q4_probed_pairs_per_step']))
    except Exception as e: print(f,'ERR',e)
P
