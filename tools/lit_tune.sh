#!/bin/bash
# chunking of the column-parallel literal kernel: repeat-rich synthetic (m = 3000) and NEAT1 x MEG3 regions
mkdir -p gpurun_out/rd && cp tests/golden/data/*.fa tests/golden/data/MEG3-DNAseq.fa.gz gpurun_out/rd/ && mkdir -p gpurun_out/rd/out
for ch in 0 4 8 16 32; do
  echo "== LTG_LIT_CH=$ch"
  LTG_LIT_CH=$ch timeout 300 python tools/repeat_probe.py 5 30 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('repeat', j['seconds'], j['gcups'])"
  for rna in H19 NEAT1; do
    ( cd gpurun_out/rd && LTG_LIT_CH=$ch ../../fasim-longtarget_b200/fasim -f1 MEG3-DNAseq.fa.gz -f2 $rna.fa -O out/ -lg 60 | grep -E "b200" | sed -e 's/.*gpu_scan_ms/gpu_scan_ms/' | tr '\n' ' '; echo " [$rna]" )
  done
done
