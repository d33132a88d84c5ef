#!/bin/bash
# Same-box A/B with a correctness gate: per-layer check of every variant, then alternating bench runs.
# usage: gpu_ab2.sh libA.so libB.so ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for lib in "$@"; do
  WATERNET_B200_LIB="$PWD/waternet_b200/$lib" timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check_$lib.log 2>&1
  echo "$lib layer_check exit $?"; grep -E "final|FAILED|bad" gpurun_out/layer_check_$lib.log | awk '$0 ~ /FAILED/ || /final/' | tail -4
done
bash tools/gpu_ab.sh "$@"
