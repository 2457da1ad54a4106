#!/bin/bash
# One GPU-box visit: parity tests, bench lines (both arms), launch list, full ncu capture of one stage-i pass.
# usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
if [ "${2:-tests}" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/${TAG}_pytest_gpu.log
  tail -3 $O/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_n1_reference_arm.json 2> $O/${TAG}_ref.err; echo "ref rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config3 --no-e2e > $O/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc $?"
python tools/stage1_once.py 10000000 3 > $O/${TAG}_s1.log 2>&1; echo "s1 rc $?"; tail -2 $O/${TAG}_s1.log
PER=$(grep -o 'per pass: [0-9]*' $O/${TAG}_s1.log | grep -o '[0-9]*$'); PRE=$(grep -o 'before the passes: [0-9]*' $O/${TAG}_s1.log | grep -o '[0-9]*$')
if [ -n "$PER" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $((PRE + 2 * PER)) -c $PER -f -o $O/${TAG}_stage1_full \
    python tools/stage1_once.py 10000000 3 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc $?"
  ls -la $O/${TAG}_stage1_full.ncu-rep
fi
