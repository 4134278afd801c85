#!/bin/bash
# literal-task statistics on the shipped example data (all 532 MEG3 regions x each example lncRNA), with and without the Q4 probe
D=tests/golden/data
mkdir -p gpurun_out/rd && cd gpurun_out/rd && rm -rf out && mkdir out
cp ../../$D/*.fa ../../$D/MEG3-DNAseq.fa.gz .
for rna in MEG3-ENST00000451743 H19 MALAT1 NEAT1; do
  for np in 0 1; do
    LTG_NO_Q4PROBE=$np ../../fasim-longtarget_b200/fasim -f1 MEG3-DNAseq.fa.gz -f2 $rna.fa -O out/ -lg 60 | grep -E "b200|Running" | tr '\n' ' '; echo " [$rna noprobe=$np]"
  done
done
