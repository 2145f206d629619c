#!/bin/bash
# visit 17 (1 GPU): programmatic dependent launch in automatic mode (on for latency-bound QMIX steps): full GPU suite + sweeps
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu17.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep17.log 2>> gpurun_out/sweep17.err; }
run --workload qmix_3m
run --workload qmix_3m --opt pdl=0
run --workload qmix_mpe_spread
run --workload qmix_mpe_spread --opt pdl=0
run --workload qmix_2s3z
run --workload qmix_8m_per
run --workload mqmix_mpe_spread
run --workload mqmix_mpe_spread --opt pdl=0
run --workload rmaddpg_spread
run --workload rmaddpg_spread --opt pdl=1
run --workload rmatd3_spread
run --workload rmatd3_spread --opt pdl=1
cat gpurun_out/sweep17.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench17.json; tail -n 3 gpurun_out/bench17.err
echo done
