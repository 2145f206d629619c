#!/bin/bash
set -u
mkdir -p gpurun_out
for o in "mixer_rm=1" "mixer_rm=2" "front_bwd_rm=2" "front_bwd_rm=3" "front_bwd_rm=4"; do
  timeout 200 python bench.py --steps 300 --warmup 20 --buffer 1024 --opt $o > gpurun_out/sweep_$o.json 2> gpurun_out/sweep_$o.err
  python - "$o" <<'PY'
import json,sys
o=sys.argv[1]
d=json.loads([l for l in open("gpurun_out/sweep_%s.json"%o).read().splitlines() if l.startswith("{")][-1])
print(o, round(d["value"]), "steps/s", {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items() if k in ("k_mixer","k_front_bwd","k_grad_reduce","k_adam","k_front_fwd_tc","k_tc_prep_weights")})
PY
done
