#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/debug_input_grads.py 2>&1 | tee gpurun_out/debug_input_grads.log | tail -20
