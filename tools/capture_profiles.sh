#!/bin/bash
# ncu captures of the per-stage kernels of one rollup-bench segment proof (run under gpurun, one GPU).
# Each capture: second proof of `tools/prove_once.py --count 2` (the first one is the warm-up incl. table set-up).
# The .ncu-rep files are converted to text on the box (details page + tools/ncu_summary.py) and deleted: a report with
# --import-source is 20-30 MB and gpurun_out/ is capped at 64 MiB.
# usage: tools/capture_profiles.sh <tag>   -> gpurun_out/prof_<kernel>_<tag>_{details,summary}.txt
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name, kernel regex, launches to skip
  local rep=/tmp/prof_$1_$TAG
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $rep \
      python tools/prove_once.py --count 2 > $OUT/ncu_$1.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page details > $OUT/prof_$1_${TAG}_details.txt 2>&1
    python tools/ncu_summary.py $rep.ncu-rep > $OUT/prof_$1_${TAG}_summary.txt 2>&1
    rm -f $rep.ncu-rep
    head -12 $OUT/prof_$1_${TAG}_summary.txt
  else
    echo "capture of $1 failed"; tail -3 $OUT/ncu_$1.log
  fi
}
cap tc_lde4096_v2 '^k_tc_lde4096_v2' 7    # trace coset LDE (204 columns x 16 cosets) of the second proof
cap eval_constraints '^k_eval_constraints' 1
cap deep '^k_deep$' 1
cap fri_fold '^k_fri_fold' 11             # layer 0 fold of the second proof
cap tc_merkle_level '^k_tc_merkle_level' 27   # first (32768-node) level of the trace tree
