#!/bin/bash
# visit 19 (1 GPU): k_front_bwd_tc in streamed mode (two CTAs per SM, LayerNorm sums through the side array)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tc or wide or golden or mlp or late" > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu19.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep19.log 2>> gpurun_out/sweep19.err; }
run --workload qmix_8m_per
run --workload qmix_8m_per --opt front_bwd_tc_stream=0
run --workload qmix_2s3z
run --workload qmix_2s3z --opt front_bwd_tc_stream=0
run --workload qmix_3m
run --workload qmix_mpe_spread
cat gpurun_out/sweep19.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8m_v19.csv \
    python bench.py --quick --workload qmix_8m_per --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_8m_v19.log 2>&1
python tools/launch_summary.py gpurun_out/launches_8m_v19.csv
echo done
