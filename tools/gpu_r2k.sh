#!/bin/bash
# Round 2, visit K: bisect of the NEAT1 failure of visit J (bulk-copy staging vs literal-only batches vs the previous build) and
# A/B of the screened carried-F recording.
TAG=${1:-r02k}
mkdir -p gpurun_out
V=fasim-longtarget_b200/variants
N="python bench.py --config neat1 --steps 1 --warmup 0 --no-cpu-baseline"
timeout 300 $N > gpurun_out/${TAG}_neat1_new.json 2> gpurun_out/${TAG}_neat1_new.err; echo "neat1 new rc=$?"
LTG_LITONLY_OLD=1 timeout 300 $N > gpurun_out/${TAG}_neat1_new_litold.json 2> gpurun_out/${TAG}_neat1_new_litold.err; echo "neat1 new litonly-old rc=$?"
LTG_FREC=0 LTG_LITONLY_OLD=1 timeout 300 $N > gpurun_out/${TAG}_neat1_new_litold_probe.json 2> gpurun_out/${TAG}_neat1_new_litold_probe.err; echo "neat1 new litonly-old probe rc=$?"
FASIM_B200_LIB=$V/libfasim_b200_nobulk.so timeout 300 $N > gpurun_out/${TAG}_neat1_nobulk.json 2> gpurun_out/${TAG}_neat1_nobulk.err; echo "neat1 nobulk rc=$?"
FASIM_B200_LIB=$V/libfasim_b200_base.so timeout 300 $N > gpurun_out/${TAG}_neat1_base.json 2> gpurun_out/${TAG}_neat1_base.err; echo "neat1 base rc=$?"
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
timeout 600 $B > gpurun_out/${TAG}_ab_new.json 2> gpurun_out/${TAG}_ab_new.err; echo "new rc=$?"
FASIM_B200_LIB=$V/libfasim_b200_nobulk.so timeout 600 $B > gpurun_out/${TAG}_ab_nobulk.json 2> gpurun_out/${TAG}_ab_nobulk.err; echo "nobulk rc=$?"
FASIM_B200_LIB=$V/libfasim_b200_base.so timeout 600 $B > gpurun_out/${TAG}_ab_base.json 2> gpurun_out/${TAG}_ab_base.err; echo "base rc=$?"
LTG_FREC=1 timeout 600 $B > gpurun_out/${TAG}_ab_frec1.json 2> gpurun_out/${TAG}_ab_frec1.err; echo "frec1 rc=$?"
for cfg in h19 malat1; do
  timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_auto.json 2> gpurun_out/${TAG}_${cfg}_auto.err; echo "$cfg auto rc=$?"
  LTG_FREC=0 timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_probe.json 2> gpurun_out/${TAG}_${cfg}_probe.err; echo "$cfg probe rc=$?"
done
timeout 900 python -m pytest tests -m gpu -q -k "q4_probe or long_lncrnas or repeat_rich" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 12 gpurun_out/${TAG}_pytest.log | cut -c1-300
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f ms=%.1f scan_gcups=%.0f scan_ms=%.0f win=%.0f rows=%s lit=%s probed=%s'%(j['value'], j['ms_per_step'], j['roofline']['achieved'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['triplex_rows_per_step'], j['literal_tasks_per_step'], j['q4_probed_pairs_per_step']))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do echo "$f: $(tail -n 1 $f | cut -c1-300)"; done
