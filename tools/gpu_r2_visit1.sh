#!/bin/bash
# round-2 first visit: whole GPU suite (no -x), MLP diagnostic, bench, option sweeps of the never-run tensor-core kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_gpu.log
timeout 300 python tools/diag_mlp.py > gpurun_out/diag_mlp.log 2>&1; tail -n 30 gpurun_out/diag_mlp.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
for o in 0 1 2; do for w in qmix_3m qmix_mpe_spread; do
  timeout 200 python bench.py --workload $w --quick --steps 100 --warmup 10 --buffer 2000 --opt wgrad_tc=$o >> gpurun_out/sweep_wgrad.log 2>> gpurun_out/sweep_wgrad.err
done; done; cat gpurun_out/sweep_wgrad.log
for w in qmix_8m_per qmix_2s3z; do
  timeout 200 python bench.py --workload $w --quick --steps 50 --warmup 5 --buffer 2000 >> gpurun_out/sweep_wide.log 2>> gpurun_out/sweep_wide.err
  timeout 200 python bench.py --workload $w --quick --steps 50 --warmup 5 --buffer 2000 --opt front_tc_wide=1 >> gpurun_out/sweep_wide.log 2>> gpurun_out/sweep_wide.err
  for o in 1 2; do
    timeout 200 python bench.py --workload $w --quick --steps 50 --warmup 5 --buffer 2000 --opt front_tc_wide=1 --opt wgrad_tc=$o >> gpurun_out/sweep_wide.log 2>> gpurun_out/sweep_wide.err
  done
done; cat gpurun_out/sweep_wide.log; tail -n 5 gpurun_out/sweep_wide.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tcbwd.csv \
    python bench.py --quick --steps 3 --warmup 3 --buffer 512 --opt wgrad_tc=2 > gpurun_out/ncu_launch_tcbwd.log 2>&1
echo done
