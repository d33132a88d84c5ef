"""Bring-up aid: native input-image gradients vs float64 autograd, split by tensor/channel/border."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200.net import WaterNet

torch.manual_seed(0)
n, h, w = 2, 29, 43


def run(tag, mutate=None, scale=3.0):
    m = WaterNet().cuda().train()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(scale)
        if mutate:
            mutate(m)
    ins = [(torch.randint(0, 256, (n, 3, h, w)).float() / 255).cuda().requires_grad_(True) for _ in range(4)]
    target = torch.rand(n, 3, h, w).cuda()
    out = m(*ins)
    torch.nn.functional.mse_loss(out, target).backward()
    m64 = copy.deepcopy(m).double()
    ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    out64 = m64._graph(*ins64)
    torch.nn.functional.mse_loss(out64, target.double()).backward()
    print(tag, "out rel", ((out.double() - out64).abs().max() / out64.abs().max()).item())
    for name, a, b in zip("x wb he gc".split(), ins, ins64):
        g, r = a.grad.double(), b.grad
        d = g - r
        rel = (d.norm() / r.norm()).item()
        inner = (d[..., 3:-3, 3:-3].norm() / r[..., 3:-3, 3:-3].norm()).item()
        perc = [(d[:, c].norm() / r[:, c].norm()).item() for c in range(3)]
        ratio = ((g * r).sum() / (r * r).sum()).item()
        print(f"  {name}: rel {rel:.2e} interior {inner:.2e} per-channel {['%.1e' % v for v in perc]} scale {ratio:.5f}")
    errs = {}
    for (k, p), (_, q) in zip(m.named_parameters(), m64.named_parameters()):
        if q.grad is not None and q.grad.norm() > 0:
            errs[k] = ((p.grad.double() - q.grad).norm() / q.grad.norm()).item()
    print("  params:", " ".join(f"{k.replace('.weight', '.w').replace('.bias', '.b')}={v:.1e}" for k, v in errs.items()
                               if k.startswith("cmg.conv1") or k.startswith("cmg.conv2") or k.startswith("wb_refiner.conv1")),
          "worst", max(errs.values()))


for size in [(2, 29, 43), (1, 64, 96)]:
    n, h, w = size
    run(f"all {size}")
n, h, w = 2, 29, 43
run("no refiner conv1", lambda m: [getattr(m, r).conv1.weight.zero_() for r in ("wb_refiner", "ce_refiner", "gc_refiner")])
run("no cmg conv1", lambda m: m.cmg.conv1.weight.zero_())


def all_active(m):  # every ReLU of the confidence-map stack stays active: no mask flips, smooth gradient
    for i in range(1, 8):
        getattr(m.cmg, f"conv{i}").bias.fill_(2.0)


run("cmg ReLUs always active (weights x0.2, bias 2)", all_active, scale=0.2)
