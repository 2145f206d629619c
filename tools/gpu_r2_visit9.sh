#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep9.log 2>> gpurun_out/sweep9.err; }
for w in qmix_3m qmix_mpe_spread; do
  run --workload $w
  run --workload $w --opt side_prio=1
  run --workload $w --opt side_prio=-1
  run --workload $w --opt front_tc=0
  run --workload $w --opt gather_tma=0
  run --workload $w --opt front_tc_threads=128
done
run --workload qmix_8m_per
run --workload qmix_8m_per --opt gather_tma=0
run --workload qmix_2s3z
run --workload mqmix_mpe_spread
run --workload rmaddpg_spread
cat gpurun_out/sweep9.log; tail -n 5 gpurun_out/sweep9.err
timeout 200 python tools/gather_sweep.py > gpurun_out/gather_sweep_tma.log 2> gpurun_out/gather_sweep.err; tail -n 3 gpurun_out/gather_sweep.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
echo done
