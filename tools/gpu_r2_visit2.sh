#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -s 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -n "kink-aware\|passed\|failed\|FAILED" gpurun_out/pytest_gpu.log | tail -n 30
for w in qmix_3m qmix_mpe_spread qmix_8m_per qmix_2s3z; do
  timeout 200 python bench.py --workload $w --quick --steps 100 --warmup 10 --buffer 2000 >> gpurun_out/sweep2.log 2>> gpurun_out/sweep2.err
done; cat gpurun_out/sweep2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_gru_fwd|k_gru_bwd' -s 4 -c 2 \
    -o gpurun_out/prof_r02a_gru -f python bench.py --quick --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
echo done
