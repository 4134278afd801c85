#!/bin/bash
# Round 2, visit M: Q4 certification on the device (taint sweep), bulk-copy staging as the default: full parity suite, real-data
# configs with the taint sweep (default) and with the probe sweep (LTG_Q4_TAINT=0), headline A/B.
TAG=${1:-r02m}
mkdir -p gpurun_out
V=fasim-longtarget_b200/variants
timeout 900 python -m pytest tests -m gpu -q -k "taint_certification" -s > gpurun_out/${TAG}_pytest_taint.log 2>&1; echo "pytest taint rc=$?"
tail -n 15 gpurun_out/${TAG}_pytest_taint.log | cut -c1-400
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_q4_taint_certification_on_device > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 12 gpurun_out/${TAG}_pytest.log | cut -c1-300
B="python bench.py --steps 2 --warmup 1 --region-mbp 20 --no-cpu-baseline"
timeout 600 $B > gpurun_out/${TAG}_ab_new.json 2> gpurun_out/${TAG}_ab_new.err; echo "new rc=$?"
FASIM_B200_LIB=$V/libfasim_b200_nobulk.so timeout 600 $B > gpurun_out/${TAG}_ab_nobulk.json 2> gpurun_out/${TAG}_ab_nobulk.err; echo "nobulk rc=$?"
for cfg in neat1 h19 malat1 meg3; do
  timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_taint.json 2> gpurun_out/${TAG}_${cfg}_taint.err; echo "$cfg taint rc=$?"
  LTG_Q4_TAINT=0 LTG_FREC=0 timeout 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_probe.json 2> gpurun_out/${TAG}_${cfg}_probe.err; echo "$cfg probe rc=$?"
done
LTG_FREC=1 timeout 300 python bench.py --config neat1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_neat1_taint_frec.json 2> gpurun_out/${TAG}_neat1_taint_frec.err; echo "neat1 taint+frec rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_neat1.csv python bench.py --config neat1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_neat1.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02m_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value=%.0f ms=%.1f scan_gcups=%.0f scan_ms=%.0f win=%.0f rows=%s lit=%s probed=%s'%(j['value'], j['ms_per_step'], j['roofline']['achieved'], j['stage_ms_per_step']['scan_kernel'], j['stage_ms_per_step']['window'], j['triplex_rows_per_step'], j['literal_tasks_per_step'], j['q4_probed_pairs_per_step']))
    except Exception as e: print(f,'ERR',e)
P
for f in gpurun_out/${TAG}_*.err; do echo "$f: $(tail -n 1 $f | cut -c1-300)"; done
python tools/launch_summary.py gpurun_out/${TAG}_launches_neat1.csv | head -32
