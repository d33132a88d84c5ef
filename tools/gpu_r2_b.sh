#!/bin/bash
# round 2, second GPU pass: conv3 + fused conv4 tail layer -- layer check, the new tests, same-box A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WN_CHECK_MODE=bf16_fp8 timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check_fused.log 2>&1; echo "layer check exit $?"
grep -E "shape|layer  [23]|final|FAILED" gpurun_out/layer_check_fused.log | tail -24
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or vs_golden or multi_pass or enhance_u8 or cuda_graph or resize or loader or metrics or reference_loop or every_weight_set or ragged" > gpurun_out/pytest_b.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_b.log
bash tools/gpu_ab3.sh libwaternet_b200.so libwaternet_b200.so:256 libwaternet_b200_v1.so
