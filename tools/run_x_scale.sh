#!/bin/bash
# Column-sharded single-segment proof on N GPUs of one box (BASELINE.json configs[4]): NCCL correctness check, then the x-synth bench line(s).
# usage: tools/run_x_scale.sh N TAG LOGN...     (writes gpurun_out/TAG_*.json / .log); CHECK=0 skips the correctness check
N=$1; TAG=$2; shift 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
mkdir -p gpurun_out
if [ "${CHECK:-1}" = "1" ]; then
  timeout 300 $TR tools/sharded_prove_check.py --logn 12 --verify > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"
  grep "world=\|SHARDED" gpurun_out/${TAG}_check.log
fi
for L in "$@"; do
  timeout 900 $TR bench.py --gpus $N --config x-synth --x-logn $L --steps 3 --warmup 1 > gpurun_out/${TAG}_x${L}.json 2> gpurun_out/${TAG}_x${L}.err; echo "x-synth 2^$L rc=$?"
  tail -c 400 gpurun_out/${TAG}_x${L}.err | grep -v Warning
  cut -c1-330 gpurun_out/${TAG}_x${L}.json
done
