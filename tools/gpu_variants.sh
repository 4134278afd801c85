#!/bin/bash
# A/B of scan-kernel tuning variants on the GPU box: scan-kernel GCUPS (roofline.achieved) per variant
mkdir -p gpurun_out
for lib in "" $(ls fasim-longtarget_b200/variants/*.so 2>/dev/null); do
  name=${lib:-product}
  FASIM_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 2 --warmup 1 --region-mbp ${1:-5} --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('$name', 'scan_gcups=%.0f' % j['roofline']['achieved'], 'value=%.0f' % j['value'], 'rows=%d' % j['triplex_rows_per_step'], j['stage_ms_per_step'])
"
done | tee gpurun_out/variants.txt
