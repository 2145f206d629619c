#!/bin/bash
# 8-GPU box, second scaling visit: the flag-in-data exchange at N = 8 (quick lines + the full bench line the driver's scaling run produces)
set -u
mkdir -p gpurun_out
P=29700
export MARL_B200_P2P=1
timeout 200 python bench.py --gpus 1 --quick --steps 300 --warmup 20 --workload qmix_3m >> gpurun_out/scale2.log 2>> gpurun_out/scale2.err
for w in qmix_3m qmix_2s3z; do
  P=$((P+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus 8 --quick --steps 300 --warmup 20 --workload $w 2>>gpurun_out/scale2.err | tail -n 1 >> gpurun_out/scale2.log
done
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus 8 --steps 100 --warmup 10 2>>gpurun_out/scale2.err | tail -n 1 > gpurun_out/bench_8gpu.json
P=$((P+1))
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus 4 --quick --steps 300 --warmup 20 --workload qmix_3m 2>>gpurun_out/scale2.err | tail -n 1 >> gpurun_out/scale2.log
cut -c1-500 gpurun_out/scale2.log; cut -c1-600 gpurun_out/bench_8gpu.json; grep -v "double Q\|OMP_NUM\|\*\*\*\*\|NCCL version\|^$\|UserWarning\|_VF.gru" gpurun_out/scale2.err | tail -n 8
echo done
