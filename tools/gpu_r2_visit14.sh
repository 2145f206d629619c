#!/bin/bash
# visit 14: shared-memory carveout preference (co-residency of the two branches' kernels)
set -u
mkdir -p gpurun_out
run() { timeout 200 python bench.py --quick --steps 200 --warmup 20 --buffer 2000 "$@" >> gpurun_out/sweep14.log 2>> gpurun_out/sweep14.err; }
for w in qmix_3m qmix_mpe_spread qmix_2s3z qmix_8m_per mqmix_spread rmaddpg_spread rmatd3_spread; do
  run --workload $w
  run --workload $w --opt smem_carveout=-1
done
run --workload qmix_3m --opt gru_wgrad_split=0
run --workload qmix_3m --opt side_prio=1
run --workload qmix_3m --opt side_prio=-1
cat gpurun_out/sweep14.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu14.log
echo done
