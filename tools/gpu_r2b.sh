#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_parity_abs_gpu.py tests/test_sharded_gpu.py -q --maxfail=30 -s > $O/pytest_r2b1.log 2>&1; echo "pytest1 rc=$?"; tail -8 $O/pytest_r2b1.log
python -m pytest tests/test_pipeline_gpu.py -q --maxfail=30 -s -k "clip or calc or deepcache or main_py" > $O/pytest_r2b2.log 2>&1; echo "pytest2 rc=$?"; tail -8 $O/pytest_r2b2.log
python tools/parity_report.py > $O/parity_r2b.txt 2>&1; echo "parity rc=$?"; tail -40 $O/parity_r2b.txt
python bench.py --steps 3 --warmup 3 > $O/bench_r2b.json 2> $O/bench_r2b.err; echo "bench rc=$?"; tail -3 $O/bench_r2b.err; cut -c1-300 $O/bench_r2b.json
for c in deep_cache consistency_model two_schedulers; do
  python bench.py --config $c --steps 2 --warmup 3 > $O/bench_r2b_$c.json 2> $O/bench_r2b_$c.err; echo "bench $c rc=$?"; tail -3 $O/bench_r2b_$c.err; cut -c1-300 $O/bench_r2b_$c.json
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
