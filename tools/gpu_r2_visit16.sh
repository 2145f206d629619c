#!/bin/bash
# visit 16 (1 GPU): staging loads batched in k_wgrad_tc / k_front_fwd_tc_wide; ncu --set full of the 8m step's kernels; PDL re-test
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tc or wide or golden" > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu16.log
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep16.log 2>> gpurun_out/sweep16.err; }
run --workload qmix_8m_per
run --workload qmix_2s3z
run --workload qmix_3m
run --workload qmix_3m --opt pdl=1
run --workload qmix_2s3z --opt pdl=1
run --workload qmix_8m_per --opt pdl=1
run --workload mqmix_mpe_spread
run --workload rmaddpg_spread
cat gpurun_out/sweep16.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_front_fwd_tc_wide|k_gru_fwd2|k_mid|k_gru_bwd2|k_front_bwd_tc|k_wgrad_tc|k_mix_hyper_fwd|k_mix_hyper_bwd' -s 16 -c 8 \
    -o gpurun_out/prof_r02d_8m -f python bench.py --quick --workload qmix_8m_per --steps 2 --warmup 2 --buffer 512 > gpurun_out/ncu_full_8m.log 2>&1
ls -la gpurun_out/prof_r02d_8m.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8m_v16.csv \
    python bench.py --quick --workload qmix_8m_per --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch_8m_v16.log 2>&1
python tools/launch_summary.py gpurun_out/launches_8m_v16.csv
echo done
