#!/bin/bash
# ncu --set full of the second default-mode forward of one 1080p frame (8 conv launches + gather + the 10 conditional
# bf16x3 re-run launches that return at once), with both fusions on and with both off.  The reports are condensed to
# CSV on the box (gpurun copies back at most 64 MiB); only the fused report itself comes back.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|gather_sigmoid" -s 19 -c 19 -o /tmp/prof_r2_fused \
  python tools/profile_forward.py 1 1080 1920 default > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused exit $?"
WATERNET_B200_DEBUG_FLAGS=768 timeout 900 ncu --set full --clock-control none -k regex:"conv_umma|gather_sigmoid" -s 20 -c 20 -o /tmp/prof_r2_unfused \
  python tools/profile_forward.py 1 1080 1920 default > gpurun_out/ncu_unfused.log 2>&1; echo "ncu unfused exit $?"
for v in fused unfused; do
  ncu -i /tmp/prof_r2_$v.ncu-rep --page raw --csv > gpurun_out/prof_r2_${v}_raw.csv 2>/dev/null
  python tools/summarize_ncu.py full /tmp/prof_r2_$v.ncu-rep gpurun_out/r2_umma_kernels_${v}_1080p_n1.csv > /dev/null
done
python - <<'PY'
import csv
for v in ("fused", "unfused"):
    print(v)
    for r in csv.DictReader(open(f"gpurun_out/r2_umma_kernels_{v}_1080p_n1.csv")):
        if float(r["time_ms"]) > 0.02:
            print("  %-70s %7.3f ms  dram %.2f+%.2f GB  tensor %5.1f%%  smem %s KB" % (r["kernel"][:70], float(r["time_ms"]), float(r["dram_read_GB"]), float(r["dram_write_GB"]), float(r["tensor_pipe_pct"] or 0), r["smem_dyn_KB"]))
PY
cp /tmp/prof_r2_fused.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out | tail -8; du -sh gpurun_out
