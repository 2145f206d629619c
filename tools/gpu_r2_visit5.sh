#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | grep -v "double Q" > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
run() { timeout 200 python bench.py --quick --steps 100 --warmup 10 --buffer 2000 "$@" >> gpurun_out/sweep5.log 2>> gpurun_out/sweep5.err; }
for w in qmix_3m qmix_mpe_spread qmix_2s3z qmix_8m_per; do
  run --workload $w
  run --workload $w --opt gru_threads=256
  run --workload $w --opt gru_threads=128
done
run --workload mqmix_mpe_spread
run --workload rmaddpg_spread
run --workload rmaddpg_spread --opt gru_threads=256
cat gpurun_out/sweep5.log; tail -n 5 gpurun_out/sweep5.err
timeout 200 python tools/e2e_profile.py qmix_3m 300 > gpurun_out/e2e_profile.log 2>&1; head -n 3 gpurun_out/e2e_profile.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_gru_fwd|k_gru_bwd|k_mid|k_optim_fused' -s 8 -c 4 \
    -o gpurun_out/prof_r02b -f python bench.py --quick --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
echo done
