#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep6.log 2>> gpurun_out/sweep6.err; }
for w in qmix_3m qmix_8m_per qmix_2s3z; do
  run --workload $w
  run --workload $w --opt gather_tma=0
done
run --workload qmix_mpe_spread
run --workload mqmix_mpe_spread
cat gpurun_out/sweep6.log; tail -n 5 gpurun_out/sweep6.err
timeout 200 python tools/gather_sweep.py > gpurun_out/gather_sweep_tma.log 2> gpurun_out/gather_sweep.err; tail -n 6 gpurun_out/gather_sweep_tma.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 30 -c 13 \
    -o gpurun_out/prof_r02c -f python bench.py --quick --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
echo done
