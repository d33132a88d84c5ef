#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/bwd_check.py > gpurun_out/bwd_check.log 2>&1; echo "bwd_check exit $?"
cat gpurun_out/bwd_check.log | tail -80
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
