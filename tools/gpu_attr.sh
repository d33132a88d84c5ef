#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/umma_layer_check.py 2>&1 | grep -E "final|FAILED"
timeout 900 python tools/pipeline_attribution.py > gpurun_out/attribution.log 2>&1; echo "exit $?"; tail -7 gpurun_out/attribution.log
