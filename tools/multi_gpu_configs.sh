#!/bin/bash
# BASELINE configs 3 / 4 / 5 (and the default config 2) under torchrun on N GPUs of one box.
# Usage (under gpurun --gpus N):  bash tools/multi_gpu_configs.sh N "deep_cache consistency_model two_schedulers"
set -u
N=${1:-8}
CONFIGS=${2:-"deep_cache consistency_model two_schedulers"}
O=gpurun_out
mkdir -p $O
PORT=29611
for c in $CONFIGS; do
  PORT=$((PORT + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --config $c --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_${c}_n$N.json 2> $O/bench_r2_${c}_n$N.err
  echo "$c n=$N rc=$?"; tail -2 $O/bench_r2_${c}_n$N.err | cut -c1-300; cut -c1-400 $O/bench_r2_${c}_n$N.json
done
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv | head -9
