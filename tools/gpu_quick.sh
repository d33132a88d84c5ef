cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "peer_out or multi_pass or enhance_u8 or submodules or range or pipeline" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gg.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_gg.json") if l.startswith("{")][-1])
print("img/s %.2f e2e %.2f clk %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]), d["kernel_ms_per_step"])
PY
