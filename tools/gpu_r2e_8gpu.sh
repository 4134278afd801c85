#!/bin/bash
# Round 2, 8-GPU visit: BASELINE.json configs[4] at its stated size (64 lncRNAs x 250 Mbp) through bench.py under torchrun (one
# rank per GPU, jobs pulled from one atomic queue, HOST buffers, rows gathered on rank 0: the e2e pass only, a pass takes ~2 min)
# and, at 64 x 100 Mbp, through the drop-in binary (`fasim --queries --devices all`, one process, per-device timing breakdown).
TAG=${1:-r02e}
QN=${2:-64}
MBP=${3:-250}
CLI_MBP=${4:-100}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt 2>&1; nproc >> gpurun_out/${TAG}_gpus.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 \
    --queries $QN --region-mbp $MBP --steps 1 --warmup 3 --e2e-only > gpurun_out/${TAG}_bench_mq_n8.json 2> gpurun_out/${TAG}_bench_mq_n8.err; echo "bench mq n8 rc=$?"
tail -n 3 gpurun_out/${TAG}_bench_mq_n8.err | cut -c1-400
timeout 300 python tools/multiquery_cli.py $QN $CLI_MBP all > gpurun_out/${TAG}_cli_mq_n8.log 2>&1; echo "cli rc=$?"
tail -n 14 gpurun_out/${TAG}_cli_mq_n8.log | cut -c1-400
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench*n8.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.0f e2e=%.0f ms=%.1f gathered=%s jobs=%s'%(j['value'], j['e2e']['value'], j['ms_per_step'], j.get('rows_gathered_on_rank0_per_step'), j.get('jobs_per_step')))
    except Exception as e: print(f, 'ERR', e)
P
