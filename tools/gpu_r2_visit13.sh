#!/bin/bash
# visit 13: k_gru_wgrad beside k_front_bwd (option gru_wgrad_split), forked branch at every size
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "qmix or golden or graph" > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu13.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep13.log 2>> gpurun_out/sweep13.err; }
run --workload qmix_3m
run --workload qmix_3m --opt gru_wgrad_split=0
run --workload qmix_3m --opt side_prio=1
run --workload qmix_3m --opt side_prio=-1
run --workload qmix_mpe_spread
run --workload qmix_mpe_spread --opt gru_wgrad_split=0
run --workload qmix_2s3z
run --workload qmix_2s3z --opt wgrad_tc=0
run --workload qmix_2s3z --opt wgrad_tc=0 --opt gru_wgrad_split=0
run --workload qmix_8m_per
run --workload qmix_8m_per --opt wgrad_tc=0
run --workload mqmix_spread
cat gpurun_out/sweep13.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_3m_v13.csv \
    python bench.py --quick --workload qmix_3m --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_3m_v13.log 2>&1
python tools/launch_summary.py gpurun_out/launches_3m_v13.csv
echo done
