#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep3.log 2>> gpurun_out/sweep3.err; }
for w in qmix_3m qmix_8m_per qmix_2s3z; do
  run --workload $w
  run --workload $w --opt optim_fused=0
  run --workload $w --opt gru_fwd_rpc=2
  run --workload $w --opt gru_bwd_rpc=2
  run --workload $w --opt gru_fwd_rpc=2 --opt gru_bwd_rpc=2
done
run --workload qmix_mpe_spread
run --workload mqmix_mpe_spread
cat gpurun_out/sweep3.log; tail -n 5 gpurun_out/sweep3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
echo done
