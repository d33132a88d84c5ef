#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WN_CHECK_MODE=bf16_fp8 timeout 60 python tools/umma_layer_check.py 2>&1 | grep -E "final|FAILED" | tr '\n' ' '; echo
timeout 90 python -m pytest tests -m gpu -q -x -k "fp8_mode or cuda_graph or enhance_u8 or vs_golden" 2>&1 | tail -4
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', d['value'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'])"
