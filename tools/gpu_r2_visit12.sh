#!/bin/bash
# visit 12: 8-warp k_mid for the large SMAC shapes; split mixer / overlap policy for 8m and 2s3z
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "qmix or golden" > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu12.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep12.log 2>> gpurun_out/sweep12.err; }
for w in qmix_8m_per qmix_2s3z; do
  run --workload $w
  run --workload $w --opt mixer_split=2
  run --workload $w --opt overlap=2
  run --workload $w --opt mixer_split=2 --opt mid_fused=0
  run --workload $w --opt mixer_split=0
done
run --workload qmix_3m
run --workload qmix_mpe_spread
cat gpurun_out/sweep12.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8m_split.csv \
    python bench.py --quick --workload qmix_8m_per --opt mixer_split=2 --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_8m_split.log 2>&1
echo done
