#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WN_CHECK_MODE=bf16_fp8 timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check_kp.log 2>&1; echo "layer check exit $?"
grep -E "shape|layer  [08] |final|FAILED" gpurun_out/layer_check_kp.log | tail -16
timeout 900 python -m pytest tests -m gpu -q -x -k "k_packed or fused or submodules or recomputes or multi_pass or enhance_u8 or vs_golden or ragged or channels_last or first_layer or 1080p or cuda_graph" > gpurun_out/pytest_e.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_e.log
bash tools/gpu_ab3.sh "$@"
