#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernels.
# Usage (from the build container):  gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [quick]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
python bench.py --workload qmix_8m_per --steps 100 --warmup 10 --buffer 2000 > gpurun_out/bench_8m.json 2> gpurun_out/bench_8m.err; cat gpurun_out/bench_8m.json; tail -n 3 gpurun_out/bench_8m.err
if [ "${1:-}" != "quick" ]; then
  python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_front_bwd|k_gru_fwd|k_gru_bwd|k_mixer|k_front_fwd|k_gather' -s 20 -c 12 \
      -o gpurun_out/prof_r01 -f python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
  timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1; tail -n 5 gpurun_out/memcheck.log
  timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/racecheck.log 2>&1; tail -n 5 gpurun_out/racecheck.log
fi
ls -la gpurun_out
