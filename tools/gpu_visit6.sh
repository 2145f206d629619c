#!/bin/bash
# GPU visit 6: split mixer + forked branches -- parity, then the option sweep, then the full bench line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
for o in "overlap=1" "overlap=0" "mixer_split=0" "mixer_split_rm=2" "mixer_split_rm=2 --opt overlap=0"; do
  timeout 200 python bench.py --quick --steps 300 --warmup 20 --buffer 1024 --opt $o 2>gpurun_out/q.err | tail -n 1; tail -n 2 gpurun_out/q.err
done > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
timeout 200 python bench.py --quick --workload qmix_8m_per --steps 50 --warmup 5 --buffer 2000 2>&1 | tail -n 1
timeout 200 python bench.py --quick --workload qmix_8m_per --steps 50 --warmup 5 --buffer 2000 --opt overlap=0 2>&1 | tail -n 1
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
