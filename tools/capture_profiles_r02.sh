#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU): the two kernels the bench line's roofline objects describe, plus the launch list
# of one bench step. Each --set full capture targets the second proof of `tools/prove_once.py --count 2` (the first is the warm-up).
# The .ncu-rep files are converted to text on the box (details page + tools/ncu_summary.py) and deleted (20-30 MB each).
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name, kernel regex, launches to skip
  local rep=/tmp/prof_$1_$TAG
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $rep \
      python tools/prove_once.py --count 2 > $OUT/ncu_$1.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page details > $OUT/prof_$1_${TAG}_details.txt 2>&1
    python tools/ncu_summary.py $rep.ncu-rep > $OUT/prof_$1_${TAG}_summary.txt 2>&1
    rm -f $rep.ncu-rep
    head -20 $OUT/prof_$1_${TAG}_summary.txt
  else
    echo "capture of $1 failed"; tail -3 $OUT/ncu_$1.log
  fi
}
cap tc_hash_rows '^k_tc_hash_rows' 2       # trace row hashing of the second proof (launches: trace + constraint rows per proof)
cap tc_lde4096_v2 '^k_tc_lde4096_v2' 7     # trace coset LDE (204 columns x 16 cosets) of the second proof
# launch list of one bench step (8 segment streams; ncu serialises the launches: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $OUT/launches_${TAG}_bench.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log
