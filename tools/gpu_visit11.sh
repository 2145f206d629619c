#!/bin/bash
# GPU visit 11 (last of round 1): full parity after the episode-stride fix, final bench line
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 380 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
