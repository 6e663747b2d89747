#!/bin/bash
# development aid: build variants on the GPU box and run a pytest selection on each. usage: tools/ab_test.sh "<pytest args>" "name:flags" ...
sel="$1"; shift
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  CRL_NVCC_EXTRA="$flags" python cleanrl.jl_b200/build.py --force > /dev/null || { echo "$name: build failed"; continue; }
  echo "== $name"; python -m pytest $sel -m gpu -x -q 2>&1 | tail -25 | grep -E "passed|failed|Error|error|assert|Mismatch|Max|mismatch" | head -12
done
