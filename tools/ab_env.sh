#!/bin/bash
# development aid: bench the SAME build under different environment switches (N=1, BASELINE configs[1]).
# usage: tools/ab_env.sh "name1:VAR=1 VAR2=1" "name2:" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --no-secondary --no-cpu-baseline --steps 200 --warmup 5 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/ab_%s.json" % n) if l.startswith("{")][-1])
    k = d["kernels"]
    print("%-14s ms/update %.4f  loss_grad %.4f  rollout %.4f  gae %.4f  other %.4f  launches %s  value %.1f M  e2e %.1f M" % (
        n, d["ms_per_step"], k["loss_grad"]["ms_per_update"], k["rollout"]["ms_per_update"], k["gae"]["ms_per_update"],
        k["other"]["ms_per_update"], d.get("gpu_launches"), d["value"] / 1e6, d["e2e"]["value"] / 1e6))
except Exception as e:
    print(n, "failed", e)
PY
done
