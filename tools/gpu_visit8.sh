#!/bin/bash
# GPU visit 8 (1 GPU): deterministic step + checkpoint, gather v2 sweep, draw fast path, bench line, ncu of the new kernels
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 200 python tools/gather_sweep.py > gpurun_out/gather_sweep.log 2>gpurun_out/gather_sweep.err; cat gpurun_out/gather_sweep.log | cut -c1-250; tail -n 2 gpurun_out/gather_sweep.err
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 200 python bench.py --workload qmix_8m_per --steps 50 --warmup 5 --buffer 2000 > gpurun_out/bench_8m.json 2> gpurun_out/bench_8m.err; cut -c1-200 gpurun_out/bench_8m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_front_bwd|k_gru_fwd|k_gru_bwd|k_mix|k_front_fwd|k_gather|k_adam|k_qhead|k_grad_reduce|k_draw' -s 45 -c 26 \
    -o gpurun_out/prof_r01m -f python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | head -30
