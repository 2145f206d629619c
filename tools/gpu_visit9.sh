#!/bin/bash
# GPU visit 9: k_mid, qhead_bwd v3, zero-copy rollout stepper, host-path trims -- parity, sweep, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
{
for o in "mid_fused=1" "mid_fused=0"; do
  timeout 200 python bench.py --quick --steps 400 --warmup 20 --buffer 1024 --opt $o 2>gpurun_out/q.err; tail -n 2 gpurun_out/q.err | grep -v "double Q"
done
} > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
timeout 120 python tools/rollout_bench.py > gpurun_out/rollout_bench.log 2>&1; tail -n 2 gpurun_out/rollout_bench.log | cut -c1-400
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
