"""Bring-up aid: native backward (wn_forward_train / wn_backward) vs float64 autograd, per parameter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import forward as ofw
from test_gpu_parity import _fp64_grads, _inputs_from_rgb, _model

for shape in [(1, 16, 16), (2, 24, 40)]:
    n, h, w = shape
    torch.manual_seed(0)
    sd = ofw.synthetic_state_dict(5, 3.0)
    m = _model(5, 3.0, "default").train()
    ins = _inputs_from_rgb([ofw.synthetic_image(30 + i, h, w, "smooth") for i in range(n)])
    target = torch.rand(n, 3, h, w)
    out = m(*[t.cuda() for t in ins])
    torch.nn.functional.mse_loss(out, target.cuda()).backward()
    torch.cuda.synchronize()
    ref_out, ref = _fp64_grads(sd, ins, target)
    print("shape", shape, "forward rel err", ((out.detach().cpu().double() - ref_out).abs().max() / ref_out.abs().max()).item())
    for name, p in m.named_parameters():
        g, r = p.grad.double().cpu(), ref[name]
        rel = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
        print(f"  {name:28s} |ref|={r.norm().item():.3e} rel err={rel:.3e} {'' if rel < 2e-3 else '<<<<'}")
