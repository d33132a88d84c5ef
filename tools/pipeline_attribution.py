"""Attribute per-layer time of the tensor-core forward to pipeline pieces by switching them off
(wn_debug_set_flags).  Prints per-layer ms for one 1080p image under each flag combination."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet

VARIANTS = [(0, "normal"), (1, "epilogue: arithmetic but no stores"), (128, "epilogue: shared-memory stores instead"),
            (0, "normal (again)"), (128, "epilogue: shared-memory stores instead (again)")]

if len(sys.argv) < 2:  # driver: one subprocess per variant (a trap in one must not take the others down)
    import subprocess
    for flags, label in VARIANTS:
        res = subprocess.run([sys.executable, __file__, str(flags)], capture_output=True, text=True, timeout=300)
        line = [l for l in res.stdout.splitlines() if l.startswith("ms ")]
        print(f"{label:48s}", line[0][3:] if line else "FAILED: " + res.stderr.strip().splitlines()[-1][:120], flush=True)
    sys.exit(0)

flags = int(sys.argv[1])
torch.manual_seed(0)
eng = get_engine("cuda:0")
m = WaterNet().cuda().eval()
eng.pack_weights(m._ordered_params())
n, h, w = 2, 1080, 1920
rgb = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
pre = eng.preprocess(rgb)
ins = [pre[k] for k in ("x", "wb", "he", "gc")]
out = torch.empty((n, 3, h, w), device="cuda")
names = ["L1", "c2", "c3", "c4", "c5", "c6", "c7", "c8", "-", "r2", "r3"]
eng.lib.wn_debug_set_flags(eng.handle, flags)
for _ in range(2):
    eng.forward(*ins, mode=_lib.MODE_BF16X3, out=out)
import subprocess, statistics
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                       stdout=subprocess.PIPE, text=True)
ITERS = 20
eng.enable_timing(True)
for _ in range(ITERS):
    eng.forward(*ins, mode=_lib.MODE_BF16X3, out=out)
ms, cnt = eng.read_timings()
smi.terminate()
samples = [l.split(",") for l in smi.communicate()[0].splitlines() if "," in l]
clk = [float(a) for a, b in samples if float(b) > 300] or [0]
pw = [float(b) for a, b in samples if float(b) > 300] or [0]
r = {names[i]: round(ms[i] / ITERS / n, 3) for i in range(11) if cnt[i]}
print("ms", " ".join(f"{k}={v:6.3f}" for k, v in r.items()), " total=%.2f" % sum(r.values()),
      " clk=%.0f MHz power=%.0f W (%d samples)" % (statistics.median(clk), statistics.median(pw), len(clk)))
