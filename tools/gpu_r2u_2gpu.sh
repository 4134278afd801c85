#!/bin/bash
# Round 2, visit U (2 GPUs): the driver's launch line for N > 1 on the final bench.py (headline mode and the reference arm).
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 2 --warmup 3 --region-mbp 20 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_n2.json 2> gpurun_out/${TAG}_bench_reference_n2.err; echo "reference n2 rc=$?"
python - <<'P'
import json
for f in ['gpurun_out/r02u_bench_n2.json','gpurun_out/r02u_bench_reference_n2.json']:
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:j.get(k) for k in ['impl','value','n_gpus','ms_per_step','rows_gathered_on_rank0_per_step','triplex_rows_per_step']}, j.get('e2e'))
    except Exception as e: print(f,'ERR',e)
P
tail -n 3 gpurun_out/${TAG}_bench_n2.err | cut -c1-300
