#!/bin/bash
# final evidence visit of round 2 (1 GPU): full GPU suite, smoke, both bench arms, ncu launch list + --set full of the 12-launch 3m step,
# quick lines of the other BASELINE workloads
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu_final.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -n 1 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; cut -c1-300 gpurun_out/bench_ref_final.json
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 "$@" >> gpurun_out/sweep_final.log 2>> gpurun_out/sweep_final.err; }
run --workload qmix_3m
run --workload qmix_mpe_spread
run --workload qmix_2s3z
run --workload qmix_8m_per
run --workload mqmix_mpe_spread
run --workload rmaddpg_spread
run --workload rmatd3_spread
run --workload rmaddpg_spread_disc
run --workload rmatd3_spread_disc
cat gpurun_out/sweep_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_final.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 30 -c 14 \
    -o gpurun_out/prof_r02e -f python bench.py --quick --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full_final.log 2>&1
ls -la gpurun_out/prof_r02e.ncu-rep
echo done
