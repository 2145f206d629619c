#!/bin/bash
# visit 18 (1 GPU): k_front_fwd_tc_wide2 (weights streamed, two CTAs per SM) vs k_front_fwd_tc_wide; PDL default for the R-MADDPG update
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tc or wide or golden or maddpg or matd3" > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu18.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep18.log 2>> gpurun_out/sweep18.err; }
run --workload qmix_8m_per
run --workload qmix_8m_per --opt front_tc_wide2=0
run --workload qmix_2s3z
run --workload qmix_2s3z --opt front_tc_wide2=0
run --workload rmaddpg_spread
run --workload rmatd3_spread
run --workload rmaddpg_spread_disc
run --workload rmatd3_spread_disc
cat gpurun_out/sweep18.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8m_v18.csv \
    python bench.py --quick --workload qmix_8m_per --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_8m_v18.log 2>&1
python tools/launch_summary.py gpurun_out/launches_8m_v18.csv
echo done
