#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "input_image or empty_batch or 4k" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_new.log
bash tools/gpu_ab2.sh libwaternet_b200.so libwaternet_b200_v1.so libwaternet_b200_v2.so libwaternet_b200_v3.so
