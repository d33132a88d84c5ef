#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/pipeline_attribution.py > gpurun_out/attribution.log 2>&1; echo "exit $?"; tail -8 gpurun_out/attribution.log
