#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 300 python bench.py --steps 300 --warmup 20 --opt pdl=0 > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench nopdl exit $?"; cut -c1-260 gpurun_out/bench_nopdl.json; tail -n 3 gpurun_out/bench_nopdl.err
for w in rmaddpg_spread rmatd3_spread_disc; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --buffer 1024 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; cut -c1-260 gpurun_out/bench_$w.json; tail -n 3 gpurun_out/bench_$w.err
done
timeout 300 python bench.py --workload qmix_8m_per --steps 100 --warmup 10 --buffer 2000 > gpurun_out/bench_8m.json 2> gpurun_out/bench_8m.err; cut -c1-260 gpurun_out/bench_8m.json; tail -n 3 gpurun_out/bench_8m.err
