#!/bin/bash
# visit 21 (2 GPUs): the full bench line as the driver's scaling run launches it, on the final code (full-size replay shard per rank)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 2 --steps 100 --warmup 10 2> gpurun_out/bench_2gpu_final.err | tail -n 1 > gpurun_out/bench_2gpu_final.json
echo "exit $?"; cut -c1-700 gpurun_out/bench_2gpu_final.json; grep -v "double Q\|OMP_NUM\|\*\*\*\*\|NCCL version\|^$\|UserWarning\|_VF.gru" gpurun_out/bench_2gpu_final.err | tail -n 6
echo done
