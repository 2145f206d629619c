#!/bin/bash
# 2-GPU visit: peer-memory gradient exchange vs NCCL -- parity test, then throughput both ways
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_dp.py -q -x --timeout 380 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dp.log
tail -n 25 gpurun_out/pytest_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
MARL_B200_P2P=1 timeout 200 $TR --master-port 29711 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 2>gpurun_out/dp_p2p.err | tail -n 1
MARL_B200_P2P=0 timeout 200 $TR --master-port 29712 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 2>gpurun_out/dp_nccl.err | tail -n 1
} > gpurun_out/dp_sweep.log 2>&1
cat gpurun_out/dp_sweep.log; tail -n 4 gpurun_out/dp_p2p.err; tail -n 3 gpurun_out/dp_nccl.err
timeout 300 $TR --master-port 29713 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"; cut -c1-400 gpurun_out/bench_2gpu.json; tail -n 3 gpurun_out/bench_2gpu.err
