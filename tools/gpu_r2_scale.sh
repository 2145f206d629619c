#!/bin/bash
# 8-GPU box: weak-scaling curve of the learner (batch 32 per GPU), 3m and 2s3z shapes, exchange breakdown per rank
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
P=29800
one() {   # workload n [env]
  local w=$1 n=$2
  if [ "$n" = 1 ]; then
    timeout 200 python bench.py --gpus 1 --quick --steps 300 --warmup 20 --workload $w >> gpurun_out/scale.log 2>> gpurun_out/scale.err
  else
    P=$((P+1))
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
        bench.py --gpus $n --quick --steps 300 --warmup 20 --workload $w 2>>gpurun_out/scale.err | tail -n 1 >> gpurun_out/scale.log
  fi
}
export MARL_B200_P2P=1
one qmix_3m 1; one qmix_3m 8; one qmix_3m 4; one qmix_3m 2
one qmix_2s3z 1; one qmix_2s3z 8; one qmix_2s3z 4; one qmix_2s3z 2
export MARL_B200_P2P=0
P=$((P+1))
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus 8 --quick --steps 300 --warmup 20 2>>gpurun_out/scale.err | tail -n 1 >> gpurun_out/scale_nccl8.log
cat gpurun_out/scale.log | cut -c1-600; cat gpurun_out/scale_nccl8.log | cut -c1-300; tail -n 5 gpurun_out/scale.err
echo done
