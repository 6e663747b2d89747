#!/bin/bash
# development aid: build loss_grad_tc variants on the GPU box and bench each (N=1, BASELINE configs[1]).
# usage: tools/ab_bench.sh "name1:flags1" "name2:flags2" ...   (flags = extra nvcc -D switches; env assignments before '|')
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; rest="${spec#*:}"
  envs=""; flags="$rest"
  if [[ "$rest" == *"|"* ]]; then envs="${rest%%|*}"; flags="${rest#*|}"; fi
  CRL_NVCC_EXTRA="$flags" python cleanrl.jl_b200/build.py --force > /dev/null || { echo "$name: build failed"; continue; }
  env $envs python bench.py --no-secondary --no-cpu-baseline --steps 100 --warmup 5 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/ab_%s.json" % n) if l.startswith("{")][-1])
    k = d["kernels"]
    print("%-14s ms/update %.4f  loss_grad %.4f  rollout %.4f  value %.1f M" % (n, d["ms_per_step"], k["loss_grad"]["ms_per_update"], k["rollout"]["ms_per_update"], d["value"] / 1e6))
except Exception as e:
    print(n, "failed", e)
PY
done
