#!/bin/bash
# fp8-correction scheme: per-layer check against the fp32 path, then same-box bench of the tensor modes / variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/f8_ab.log
for lib in libwaternet_b200.so "$@"; do
  WATERNET_B200_LIB="$PWD/waternet_b200/$lib" WN_CHECK_MODE=bf16_fp8 timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check_f8_$lib.log 2>&1; echo "$lib layer_check f8 exit $?"
  grep -E "shape|layer  7|final|FAILED" gpurun_out/layer_check_f8_$lib.log | tail -16
done
run() {
WATERNET_B200_LIB="$PWD/waternet_b200/$1" timeout 300 python bench.py --mode $2 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k=d['kernel_ms_per_step']
print('$1 $2', 'img/s %.2f' % d['value'], 'clk', d['clocks']['sm_mhz'], ' '.join('%s=%.1f' % (a.split('.')[-1][:8], b) for a,b in k.items() if b > 3))
" | tee -a gpurun_out/f8_ab.log
}
for round in 1 2; do
  run libwaternet_b200.so bf16x3
  for lib in libwaternet_b200.so "$@"; do run $lib bf16_fp8; done
done
