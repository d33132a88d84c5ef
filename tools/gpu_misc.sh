#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "loader or backward or training" 2>&1 | tail -3
timeout 1800 bash tests/cli_smoke.sh > gpurun_out/cli_smoke.log 2>&1; echo "cli exit $?"; tail -14 gpurun_out/cli_smoke.log
