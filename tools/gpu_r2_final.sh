#!/bin/bash
# round-2 validation + evidence: whole gpu test-suite, smoke(), bench lines (default, bf16x3, reference arm),
# ncu --set full of one default-mode forward (condensed on the box), launch list of a short bench, sanitizer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log; grep -h "max rel err" gpurun_out/pytest_gpu.log | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --mode bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench bf16x3 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?"
python - <<'PY'
import json
for f in ("default", "bf16x3", "reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{f}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "ms", d.get("ms_per_step"), d.get("clocks"))
    if f == "default": print(d.get("kernel_ms_per_step")); print(d.get("parity")); print(d["roofline"]["frac"], d["roofline"]["achieved"], d["gpu_launches"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|gather_" -s 19 -c 19 -o /tmp/prof_r2_fwd \
  python tools/profile_forward.py 1 1080 1920 default > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/summarize_ncu.py full /tmp/prof_r2_fwd.ncu-rep gpurun_out/r2_umma_kernels_1080p_n1.csv > /dev/null
timeout 600 ncu --set full --clock-control none -k regex:"apply_kernel|stats_kernel|luts_kernel" -s 3 -c 3 -o /tmp/prof_r2_pre \
  python tools/profile_forward.py 4 1080 1920 default > gpurun_out/ncu_pre.log 2>&1; echo "ncu pre exit $?"
python tools/summarize_ncu.py full /tmp/prof_r2_pre.ncu-rep gpurun_out/r2_preprocess_kernels_1080p_n4.csv > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"; grep -c "Race reported" gpurun_out/sanitize_racecheck.log; tail -3 gpurun_out/sanitize_racecheck.log
du -sh gpurun_out
