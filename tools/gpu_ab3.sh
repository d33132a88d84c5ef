#!/bin/bash
# Same-box A/B of library variants and debug-flag settings: alternate short bench runs.
# usage: gpu_ab3.sh lib[:flags] ...     (flags = WATERNET_B200_DEBUG_FLAGS, e.g. 256 = conv3/conv4 unfused)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/ab.log
for round in 1 2 3; do
  for ent in "$@"; do
    lib="${ent%%:*}"; flags="0"; [[ "$ent" == *:* ]] && flags="${ent##*:}"
    WATERNET_B200_DEBUG_FLAGS="$flags" WATERNET_B200_LIB="$PWD/waternet_b200/$lib" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k=d['kernel_ms_per_step']
print('$ent', 'round $round', 'img/s %.2f' % d['value'], 'e2e %.2f' % d['e2e']['value'], 'clk', d['clocks']['sm_mhz'], ' '.join('%s=%.1f' % (a.split('.')[-1][:8], b) for a,b in k.items() if b > 3))
" | tee -a gpurun_out/ab.log
  done
done
