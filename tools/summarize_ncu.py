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


# the ten tensor-core launches of one forward pass, in launch order (conv_umma.cu: umma_forward_layers)
FORWARD_LAUNCHES = ["cmg.conv1+refiner.conv1x3", "cmg.conv2", "cmg.conv3", "cmg.conv4", "cmg.conv5", "cmg.conv6",
                    "cmg.conv7", "cmg.conv8", "refiner.conv2x3", "refiner.conv3x3+gate"]


def traffic(kernels_csv, out):
    """DRAM bytes per launch of each forward kernel (one 1080p image per launch) -> the JSON bench.py reads
    for roofline.traffic.  Input: the CSV written by `full` from a capture of exactly one forward pass."""
    import json
    rows = list(csv.DictReader(open(kernels_csv)))
    assert len(rows) == len(FORWARD_LAUNCHES), f"expected one forward pass ({len(FORWARD_LAUNCHES)} launches), got {len(rows)}"
    doc = {"source": f"ncu --set full, one 1920x1080 image per launch ({kernels_csv}): "
                     "dram__bytes_read.sum + dram__bytes_write.sum",
           "height": 1080, "width": 1920, "kernels": {}}
    for name, r in zip(FORWARD_LAUNCHES, rows):
        doc["kernels"][name] = {"dram_bytes_per_image": (float(r["dram_read_GB"]) + float(r["dram_write_GB"])) * 1e9,
                                "kernel": r["kernel"]}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
