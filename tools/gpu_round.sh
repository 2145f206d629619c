#!/bin/bash
# One standard GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture, summaries.
# Usage (from the build container):  gpurun --timeout 2700 -- 'bash tools/gpu_round.sh <tag> [quick]'   (~30 min with the option sweeps)
#   then:  python tools/ncu_summary.py <tag> gpurun_out/launches.csv gpurun_out/prof_<tag>.ncu-rep   (writes profiles/<tag>_ncu_*)
set -u
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
if [ "${2:-}" != "quick" ]; then
  timeout 200 python bench.py --workload qmix_8m_per --steps 50 --warmup 5 --buffer 2000 > gpurun_out/bench_8m.json 2> gpurun_out/bench_8m.err; cut -c1-200 gpurun_out/bench_8m.json
  timeout 200 python bench.py --workload qmix_mpe_spread --steps 300 --warmup 20 > gpurun_out/bench_mpe.json 2> gpurun_out/bench_mpe.err; cut -c1-200 gpurun_out/bench_mpe.json
  timeout 300 python bench.py --workload mqmix_mpe_spread --steps 300 --warmup 20 > gpurun_out/bench_mlp.json 2> gpurun_out/bench_mlp.err; cut -c1-200 gpurun_out/bench_mlp.json
  # option candidates prepared without a GPU (off by default until timed): wide-input tcgen05 front kernel on the obs-80 workloads
  for o in 0 1; do for w in qmix_8m_per qmix_2s3z; do
    timeout 200 python bench.py --workload $w --quick --steps 50 --warmup 5 --buffer 2000 --opt front_tc_wide=$o >> gpurun_out/sweep_wide.log 2>> gpurun_out/sweep_wide.err
  done; done; cat gpurun_out/sweep_wide.log
  # tensor-core backward of the front layers (1: k_wgrad_tc beside k_front_bwd, 2: k_front_bwd_tc + k_wgrad_tc) -- input widths <= 64
  for o in 0 1 2; do for w in qmix_3m qmix_mpe_spread mqmix_mpe_spread; do
    timeout 200 python bench.py --workload $w --quick --steps 100 --warmup 10 --buffer 2000 --opt wgrad_tc=$o >> gpurun_out/sweep_wgrad.log 2>> gpurun_out/sweep_wgrad.err
  done; done; cat gpurun_out/sweep_wgrad.log
  # obs-80 workloads with every tensor-core kernel on
  for w in qmix_8m_per qmix_2s3z; do for o in 1 2; do
    timeout 200 python bench.py --workload $w --quick --steps 50 --warmup 5 --buffer 2000 --opt front_tc_wide=1 --opt wgrad_tc=$o >> gpurun_out/sweep_wide_all.log 2>> gpurun_out/sweep_wide_all.err
  done; done
  # 8m with the split mixer + k_mid on ONE stream (no forked branch at this size): k_qhead + k_qhead_bwd are 150 us of the fused-mixer step
  timeout 200 python bench.py --workload qmix_8m_per --quick --steps 50 --warmup 5 --buffer 2000 --opt front_tc_wide=1 --opt wgrad_tc=2 --opt mixer_split=2 >> gpurun_out/sweep_wide_all.log 2>> gpurun_out/sweep_wide_all.err
  cat gpurun_out/sweep_wide_all.log
  for o in 0 2; do for w in rmaddpg_spread rmatd3_spread_disc; do
    timeout 200 python bench.py --workload $w --quick --steps 100 --warmup 10 --opt front_tc_wide=1 --opt wgrad_tc=$o >> gpurun_out/sweep_maddpg.log 2>> gpurun_out/sweep_maddpg.err
  done; done; cat gpurun_out/sweep_maddpg.log
  timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
  timeout 200 python tools/gather_sweep.py > gpurun_out/gather_sweep.log 2> gpurun_out/gather_sweep.err; cut -c1-200 gpurun_out/gather_sweep.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_launch.log 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 40 -c 16 \
      -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --buffer 512 > gpurun_out/ncu_full.log 2>&1
  # first profiles of the tensor-core backward and the wide forward (written without a GPU): launch list + full set
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tcbwd.csv \
      python bench.py --quick --steps 3 --warmup 3 --buffer 512 --opt wgrad_tc=2 > gpurun_out/ncu_launch_tcbwd.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_wgrad_tc|k_front_bwd_tc' -s 6 -c 6 \
      -o gpurun_out/prof_${TAG}_tcbwd -f python bench.py --quick --steps 3 --warmup 3 --buffer 512 --opt wgrad_tc=2 > gpurun_out/ncu_full_tcbwd.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_front_fwd_tc_wide|k_wgrad_tc|k_front_bwd_tc' -s 6 -c 6 \
      -o gpurun_out/prof_${TAG}_8m_tc -f python bench.py --workload qmix_8m_per --quick --steps 2 --warmup 3 --buffer 512 --opt front_tc_wide=1 --opt wgrad_tc=2 > gpurun_out/ncu_full_8m_tc.log 2>&1
fi
ls -la gpurun_out | head -40
