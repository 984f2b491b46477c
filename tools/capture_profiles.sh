#!/bin/bash
# ncu captures of the per-stage kernels of one rollup-bench segment proof (run under gpurun, one GPU).
# Each capture: second proof of `tools/prove_once.py --count 2` (the first one is the warm-up incl. table set-up).
# usage: tools/capture_profiles.sh <tag>   -> gpurun_out/prof_<kernel>_<tag>.ncu-rep
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name, kernel regex, launches to skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $OUT/prof_$1_$TAG \
      python tools/prove_once.py --count 2 > $OUT/ncu_$1.log 2>&1
  tail -1 $OUT/ncu_$1.log
}
cap tc_lde4096 '^k_tc_lde4096' 7          # trace coset LDE (204 columns x 16 cosets) of the second proof
cap eval_constraints '^k_eval_constraints' 1
cap deep '^k_deep$' 1
cap fri_fold '^k_fri_fold' 11             # layer 0 fold of the second proof
cap tc_merkle_level '^k_tc_merkle_level' 27   # first (32768-node) level of the trace tree
