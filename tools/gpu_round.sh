#!/bin/bash
# One GPU-box visit: parity tests, bench, per-operator profile, ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [skip-ncu]
set -u
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python bench.py --steps 3 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 $O/bench_$TAG.json
python tools/profile_plan.py 16 > $O/profile_plan_$TAG.log 2>&1; head -3 $O/profile_plan_$TAG.log
python tools/parity_report.py > $O/parity_$TAG.txt 2>&1; echo "parity rc=$?"
if [ "${2:-}" != "skip-ncu" ]; then
  python tools/ncu_target.py > $O/ncu_plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
      --log-file $O/launches_$TAG.csv python tools/ncu_target.py > $O/ncu_l_$TAG.log 2>&1
  echo "launch list rc=$?"
  for k in conv_gemm_kernel:14 attention2_kernel:2 attention_kernel:2 gn_:2 ln_side:1 latent_update:1; do
    name=${k%%:*}; cnt=${k##*:}
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:$name -c $cnt -f -o $O/prof_${TAG}_$name python tools/ncu_target.py > $O/ncu_f_${TAG}_$name.log 2>&1
    echo "ncu full $name rc=$?"
  done
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
