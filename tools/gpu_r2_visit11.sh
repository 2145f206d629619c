#!/bin/bash
set -u
mkdir -p gpurun_out
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep11.log 2>> gpurun_out/sweep11.err; }
run --workload rmaddpg_spread
run --workload rmatd3_spread
run --workload rmaddpg_spread_disc
run --workload rmatd3_spread_disc
cat gpurun_out/sweep11.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8m.csv \
    python bench.py --quick --workload qmix_8m_per --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_8m.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_2s3z.csv \
    python bench.py --quick --workload qmix_2s3z --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_2s3z.log 2>&1
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
echo done
