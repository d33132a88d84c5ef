#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 ./tools/umma_rate_probe 2000 2>&1 | tee gpurun_out/rate_probe.log
echo "exit ${PIPESTATUS[0]}"
timeout 120 ./tools/umma_f8_probe 2000 2>&1 | tee gpurun_out/f8_probe.log
echo "exit ${PIPESTATUS[0]}"
