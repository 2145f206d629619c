#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep8.log 2>> gpurun_out/sweep8.err; }
for w in qmix_3m qmix_mpe_spread; do
  run --workload $w
  run --workload $w --opt hyper_late=0
  run --workload $w --opt front_bwd_mma=0
  run --workload $w --opt overlap=0
  run --workload $w --opt gather_tma=0
done
run --workload qmix_8m_per
run --workload qmix_8m_per --opt hyper_late=0
run --workload qmix_8m_per --opt wgrad_tc=0
run --workload qmix_8m_per --opt wgrad_tc=0 --opt front_bwd_mma=0
run --workload qmix_2s3z
run --workload qmix_2s3z --opt wgrad_tc=0
run --workload mqmix_mpe_spread
run --workload mqmix_mpe_spread --opt front_bwd_mma=0
run --workload rmaddpg_spread
run --workload rmaddpg_spread --opt front_bwd_mma=0
run --workload rmatd3_spread
cat gpurun_out/sweep8.log; tail -n 5 gpurun_out/sweep8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  timeout 400 python -m pytest tests/test_gpu_dp.py -q -x --timeout 380 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1; tail -n 3 gpurun_out/pytest_dp.log
  {
  MARL_B200_P2P=1 timeout 200 $TR --master-port 29711 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 2>gpurun_out/dp_p2p.err | tail -n 1
  } > gpurun_out/dp_sweep.log 2>&1
  cat gpurun_out/dp_sweep.log
fi
echo done
