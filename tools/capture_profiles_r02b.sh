#!/bin/bash
# Round-2 captures of the multi-level tensor-core transform at 2^20 rows (8 columns, blowup 16): the top cross stage and the 4096-point
# kernel in its scatter mode. Converted to text on the box, .ncu-rep deleted.
set -u
OUT=gpurun_out; mkdir -p $OUT
cap() {  # name, kernel regex, launches to skip
  local rep=/tmp/prof_$1_r02
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $rep python tools/lde_once.py --logn 20 > $OUT/ncu_$1.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page details > $OUT/prof_$1_r02_details.txt 2>&1
    python tools/ncu_summary.py $rep.ncu-rep > $OUT/prof_$1_r02_summary.txt 2>&1
    rm -f $rep.ncu-rep
    head -20 $OUT/prof_$1_r02_summary.txt
  else
    echo "capture of $1 failed"; tail -3 $OUT/ncu_$1.log
  fi
}
cap tc_cross16_2pow20 '^k_tc_cross16' 2
cap tc_lde4096_scatter_2pow20 '^k_tc_lde4096_v2' 1
