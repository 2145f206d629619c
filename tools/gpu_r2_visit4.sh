#!/bin/bash
# 2-GPU visit: in-kernel peer-memory exchange -- parity, P2P vs NCCL, bench line; plus single-GPU MADDPG sweeps and the e2e host profile
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_dp.py -q -x --timeout 380 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dp.log
tail -n 15 gpurun_out/pytest_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
MARL_B200_P2P=1 timeout 200 $TR --master-port 29711 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 2>gpurun_out/dp_p2p.err | tail -n 1
MARL_B200_P2P=1 timeout 200 $TR --master-port 29714 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 --opt optim_fused=0 2>gpurun_out/dp_p2p_unfused.err | tail -n 1
MARL_B200_P2P=0 timeout 200 $TR --master-port 29712 bench.py --gpus 2 --quick --steps 300 --warmup 20 --buffer 2048 2>gpurun_out/dp_nccl.err | tail -n 1
MARL_B200_P2P=1 timeout 200 $TR --master-port 29715 bench.py --gpus 2 --quick --steps 200 --warmup 20 --buffer 2048 --workload qmix_2s3z 2>gpurun_out/dp_2s3z.err | tail -n 1
} > gpurun_out/dp_sweep.log 2>&1
cat gpurun_out/dp_sweep.log; tail -n 4 gpurun_out/dp_p2p.err; tail -n 3 gpurun_out/dp_nccl.err
timeout 300 $TR --master-port 29713 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"; cut -c1-300 gpurun_out/bench_2gpu.json; tail -n 3 gpurun_out/bench_2gpu.err
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 "$@" >> gpurun_out/sweep4.log 2>> gpurun_out/sweep4.err; }
for w in rmaddpg_spread rmatd3_spread rmaddpg_spread_disc; do
  run --workload $w
  run --workload $w --opt wgrad_tc=0 --opt front_tc_wide=0
  run --workload $w --opt wgrad_tc=0
done
run --workload qmix_8m_per --buffer 2000
run --workload qmix_2s3z --buffer 2000
cat gpurun_out/sweep4.log; tail -n 3 gpurun_out/sweep4.err
timeout 200 python tools/e2e_profile.py qmix_3m 300 > gpurun_out/e2e_profile.log 2>&1; tail -n 12 gpurun_out/e2e_profile.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo done
