"""Condense ncu artefacts from gpurun_out/ into small tracked files under profiles/.

    python tools/summarize_ncu.py full  gpurun_out/prof.ncu-rep   profiles/r1_xxx_kernels.csv
    python tools/summarize_ncu.py list  gpurun_out/launches.csv   profiles/r1_xxx_launches.csv
    python tools/summarize_ncu.py traffic profiles/r1_umma_kernels_1080p_n1.csv profiles/r1_traffic.json
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_read_GB", "dram__bytes_read.sum"),
    ("dram_write_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_clock_GHz", "sm__cycles_elapsed.avg.per_second"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
]


def to_unit(value, unit, want):
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    if want == "ms":
        return v * scale.get(unit, 1.0)
    if want == "GB":
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1.0)
    if want == "KB":
        return v * {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(unit, 1.0)
    if want == "GHz":
        return v * {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(unit, 1.0)
    return v


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [name for name, _ in FULL_METRICS])
        for r in rows[2:]:
            line = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")]
            for name, metric in FULL_METRICS:
                if metric not in hdr:
                    line.append("")
                    continue
                i = hdr.index(metric)
                want = name.rsplit("_", 1)[-1]
                line.append(f"{to_unit(r[i], units[i], want):.4g}")
            w.writerow(line)
    print(open(out).read())


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "")
        v = to_unit(r[vi], r[ui], "ms")
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    total = sum(v for _, v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct"])
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{v:.4f}", f"{100 * v / total:.2f}"])
        w.writerow(["TOTAL", sum(n for n, _ in agg.values()), f"{total:.4f}", "100"])
    print(open(out).read())


# kernel (template signature prefix) -> layer name of bench.py's kernel_ms_per_step
LAYER_OF_KERNEL = [
    ("conv_umma_kernel<7, 16, 224", "cmg.conv1+refiner.conv1x3"),
    ("conv_umma_kernel<5, 128, 128", "cmg.conv2"),
    ("conv_umma_kernel<3, 128, 128", "cmg.conv3"),       # "+conv4" when the template's tail width (12th argument) is 64
    ("conv_umma_kernel<1, 128, 64", "cmg.conv4"),
    ("conv_umma_kernel<7, 64, 64", "cmg.conv5"),
    ("conv_umma_kernel<5, 64, 64", "cmg.conv6"),
    ("conv_umma_kernel<3, 64, 64", "cmg.conv7"),         # "+conv8(taps)" with a tail of 32 columns
    ("conv_umma_kernel<3, 64, 16", "cmg.conv8"),
    ("gather_sigmoid_kernel", "cmg.conv8.gather+sigmoid"),
    ("conv_umma_kernel<5, 96, 32", "refiner.conv2x3"),   # "+conv3x3(taps)" with a tail of 96 columns
    ("gather_gate_kernel", "refiner.conv3.gather+gate"),
    ("conv_umma_kernel<3, 96, 16", "refiner.conv3x3+gate"),
]


def layer_name(kernel):
    k = kernel.replace("wn::", "")
    for prefix, name in LAYER_OF_KERNEL:
        if k.startswith(prefix):
            args = [a.strip() for a in k[k.index("<") + 1:k.rindex(">")].split(",")] if "<" in k else []
            if name == "cmg.conv3" and len(args) >= 12 and args[11] == "64":
                return "cmg.conv3+conv4"
            if name == "cmg.conv7" and len(args) >= 12 and args[11] == "32":
                return "cmg.conv7+conv8(taps)"
            if name == "refiner.conv2x3" and len(args) >= 12 and args[11] == "96":
                return "refiner.conv2x3+conv3x3(taps)"
            return name
    return None


def traffic(kernels_csv, out):
    """DRAM bytes per launch of each forward kernel (one 1080p image per launch) -> the JSON bench.py reads for
    roofline.traffic.  Input: the CSV written by `full` from a capture of one forward pass; the conditional
    launches of the range guard's bf16x3 re-run (they return at once: < 20 us) are dropped."""
    import json
    rows = [r for r in csv.DictReader(open(kernels_csv)) if float(r["time_ms"]) > 0.02]
    doc = {"source": f"ncu --set full, one 1920x1080 image per launch ({kernels_csv}): "
                     "dram__bytes_read.sum + dram__bytes_write.sum",
           "height": 1080, "width": 1920, "kernels": {}}
    for r in rows:
        name = layer_name(r["kernel"])
        assert name is not None and name not in doc["kernels"], f"unexpected or repeated kernel {r['kernel']}"
        doc["kernels"][name] = {"dram_bytes_per_image": (float(r["dram_read_GB"]) + float(r["dram_write_GB"])) * 1e9,
                                "time_ms": float(r["time_ms"]), "tensor_pipe_pct": float(r["tensor_pipe_pct"] or 0),
                                "kernel": r["kernel"]}
    doc["total_dram_bytes_per_image"] = sum(k["dram_bytes_per_image"] for k in doc["kernels"].values())
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
