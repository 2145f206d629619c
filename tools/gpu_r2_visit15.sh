#!/bin/bash
# visit 15 (2 GPUs): flag-in-data exchange lines in k_optim_fused vs the slot + flag protocol
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/pytest_dp15.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_dp15.log
P=29900
two() {
  P=$((P+1))
  MARL_B200_P2P=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus 2 --quick --steps 300 --warmup 20 "$@" 2>>gpurun_out/dp15.err | tail -n 1 >> gpurun_out/dp15.log
}
two --workload qmix_3m
two --workload qmix_3m --opt p2p_ll=0
two --workload qmix_2s3z
cut -c1-700 gpurun_out/dp15.log; tail -n 5 gpurun_out/dp15.err
echo done
