#!/bin/bash
# a gate of quick parity tests on the first listed variants, then the same-box A/B of gpu_ab3.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for ent in "$@"; do lib="${ent%%:*}"
  WATERNET_B200_LIB="$PWD/waternet_b200/$lib" timeout 600 python -m pytest tests -m gpu -q -x -k "golden or multi_pass or k_packed or margin" > gpurun_out/pytest_e_$lib.log 2>&1; echo "$lib pytest exit $?"; tail -1 gpurun_out/pytest_e_$lib.log
done
bash tools/gpu_ab3.sh "$@"
