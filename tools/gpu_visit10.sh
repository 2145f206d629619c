#!/bin/bash
# GPU visit 10: product-configuration bench (k_mid), option sweep (mid_fused, GRU rows per CTA), ncu of the final step
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_qmix.py -m gpu -q --timeout 300 -p no:cacheprovider -k "product or branch_modes or golden" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
{
for o in "mid_fused=1" "mid_fused=0" "gru_fwd_rpc=2" "gru_bwd_rpc=2" "gru_fwd_rpc=2 --opt gru_bwd_rpc=2"; do
  timeout 200 python bench.py --quick --steps 400 --warmup 20 --buffer 1024 --opt $o 2>gpurun_out/q.err; tail -n 2 gpurun_out/q.err | grep -v "double Q"
done
} > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_mid|k_qhead_bwd|k_gather|k_draw|k_policy|k_front_bwd|k_gru_fwd' -s 30 -c 10 \
    -o gpurun_out/prof_r01n -f python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
