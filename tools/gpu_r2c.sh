#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -q --maxfail=30 > $O/pytest_r2c1.log 2>&1; echo "pytest1 rc=$?"; tail -6 $O/pytest_r2c1.log
python -m pytest tests/test_pipeline_gpu.py tests/test_parity_abs_gpu.py tests/test_fullsize_gpu.py tests/test_reference_pins_gpu.py -q --maxfail=30 -s > $O/pytest_r2c2.log 2>&1; echo "pytest2 rc=$?"; tail -8 $O/pytest_r2c2.log
python bench.py --steps 3 --warmup 3 > $O/bench_r2c.json 2> $O/bench_r2c.err; echo "bench rc=$?"; tail -3 $O/bench_r2c.err; cut -c1-300 $O/bench_r2c.json
python tools/profile_plan.py 16 > $O/profile_plan_r2c.log 2>&1; head -40 $O/profile_plan_r2c.log
python bench.py --config two_schedulers --steps 1 --warmup 3 > $O/bench_r2c_two.json 2> $O/bench_r2c_two.err; echo "bench two rc=$?"; tail -3 $O/bench_r2c_two.err; cut -c1-600 $O/bench_r2c_two.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
