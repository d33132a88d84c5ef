#!/bin/bash
# final validation of a round: gpu tests, smoke(), bench lines of the default and the bf16x3 mode, reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log; grep -h "stress weights" gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 900 python bench.py --mode bf16x3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench bf16x3 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?"
python - <<'PY'
import json
for f in ("default", "bf16x3", "reference"):
    d = json.loads([l for l in open(f"gpurun_out/bench_{f}.json") if l.startswith("{")][-1])
    print(f, "value", d["value"], "e2e", d.get("e2e", {}).get("value"), "ms", d.get("ms_per_step"), d.get("clocks"))
    if f == "default": print(d["kernel_ms_per_step"]); print(d["roofline"]["frac"], d["roofline"]["achieved"], d["gpu_launches"])
PY
