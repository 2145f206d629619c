#!/bin/bash
# 2-GPU visit: data-parallel bench line + reference arm under torchrun, both bounded by timeout.
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --buffer 1024 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "dp$N exit $?"; cat gpurun_out/bench_dp$N.json | cut -c1-600; tail -n 4 gpurun_out/bench_dp$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_dp$N.json 2> gpurun_out/bench_ref_dp$N.err
echo "ref dp$N exit $?"; cat gpurun_out/bench_ref_dp$N.json | cut -c1-400; tail -n 3 gpurun_out/bench_ref_dp$N.err
timeout 200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "qmix_5ag" 2>&1 | tail -n 4
