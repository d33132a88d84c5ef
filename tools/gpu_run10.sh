#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/probe3.log
for t in mnmajor mnmajor_k64 mnmajor_shift; do timeout 60 tools/umma_probe $t >> gpurun_out/probe3.log 2>&1; echo "[$t exit $?]" >> gpurun_out/probe3.log; done
cat gpurun_out/probe3.log
bash tools/gpu_run3.sh
