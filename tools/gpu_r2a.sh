#!/bin/bash
# Round-2 first GPU visit: full GPU test suite, smoke, absolute parity report, bench.
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_r2a.log
python -m pytest tests -m gpu -q --maxfail=12 -s > $O/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_r2a.log
python tools/parity_report.py > $O/parity_r2a.txt 2>&1; echo "parity rc=$?"; tail -30 $O/parity_r2a.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_r2a.json 2> $O/bench_r2a.err; echo "bench rc=$?"; cut -c1-400 $O/bench_r2a.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
