#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gemm_gpu.py -q --maxfail=30 > $O/pytest_r2d1.log 2>&1; echo "pytest1 rc=$?"; tail -6 $O/pytest_r2d1.log
python -m pytest tests/test_pipeline_gpu.py tests/test_parity_abs_gpu.py tests/test_fullsize_gpu.py -q --maxfail=30 -s > $O/pytest_r2d2.log 2>&1; echo "pytest2 rc=$?"; tail -8 $O/pytest_r2d2.log
python tools/profile_plan.py 16 > $O/profile_plan_r2d.log 2>&1; head -60 $O/profile_plan_r2d.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_r2d.json 2> $O/bench_r2d.err; echo "bench rc=$?"; tail -3 $O/bench_r2d.err; cut -c1-300 $O/bench_r2d.json
python tools/bench_vae.py 16 > $O/vae_r2d.log 2>&1; tail -2 $O/vae_r2d.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
