#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WN_CHECK_MODE=bf16_fp8 timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check_fused.log 2>&1; echo "layer check exit $?"
grep -E "shape|layer  [23]|final|FAILED" gpurun_out/layer_check_fused.log | tail -16
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or submodules or recomputes or vs_golden or multi_pass or enhance_u8 or cuda_graph or loader or every_weight_set or ragged or recomputes or 1080p or backward or training or gradients" > gpurun_out/pytest_c.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_c.log
bash tools/gpu_ab3.sh "$@"
