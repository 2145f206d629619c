#!/bin/bash
# visit 20 (1 GPU): two sequence rows per CTA in the 128-thread recurrences at the large shapes (option gru_rows); full GPU suite on the final code
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu20.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep20.log 2>> gpurun_out/sweep20.err; }
run --workload qmix_8m_per
run --workload qmix_8m_per --opt gru_rows=1
run --workload qmix_2s3z
run --workload qmix_2s3z --opt gru_rows=1
run --workload qmix_2s3z --opt gru_rows=2
run --workload qmix_3m
cat gpurun_out/sweep20.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_8m_v20.csv \
    python bench.py --quick --workload qmix_8m_per --steps 2 --warmup 2 --buffer 512 > gpurun_out/ncu_launch_8m_v20.log 2>&1
python tools/launch_summary.py gpurun_out/launches_8m_v20.csv
echo done
