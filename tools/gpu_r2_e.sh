#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or submodules or recomputes or multi_pass or enhance_u8 or k_packed" > gpurun_out/pytest_e.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_e.log
bash tools/gpu_ab3.sh "$@"
