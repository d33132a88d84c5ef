#!/bin/bash
# ncu evidence for profiles/: full capture of the conv kernels (one 1080p image), launch list of a
# short bench, backward kernels.  Run under gpurun; condense with tools/summarize_ncu.py.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 10 -o gpurun_out/prof_umma \
  python tools/profile_forward.py 1 1080 1920 bf16x3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|stats_kernel|luts_kernel" -s 3 -c 3 -o gpurun_out/prof_pre \
  python tools/profile_forward.py 4 1080 1920 bf16x3 > gpurun_out/ncu_pre.log 2>&1; echo "ncu pre exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bf16x3.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_umma|gate_bwd|bias_grad" -s 20 -c 14 -o gpurun_out/prof_bwd \
  python tools/bench_train.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd exit $?"
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench_train exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?"
timeout 900 python bench.py --mode fp32 --steps 2 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?"
