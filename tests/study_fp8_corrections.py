"""Numerical study (CPU emulation, not a test): how much accuracy the forward keeps when the two bf16
correction passes of the bf16x3 scheme (a_lo x w, a_hi x w_lo) are evaluated with fp8 operands instead --
e5m2 / e4m3 activations x e4m3 weights with a static per-layer power-of-two scale -- as tcgen05
kind::f8f6f4 MMAs would at twice the bf16 rate.  Max relative error of the network output against a
float64 evaluation, stress weights (gain 1 and 3), 96x96 images.  Results are quoted in DESIGN.md section 9.

    python tests/study_fp8_corrections.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import torch.nn.functional as F
from oracle import forward as ofw, preprocess as opre
torch.set_num_threads(16)
def bf16(x): return x.to(torch.bfloat16).to(torch.float64)
def q8(x, dt): return x.to(torch.float32).to(dt).to(torch.float64)
def conv_scheme(a, w, b, k, scheme):
    a=a.double(); w=w.double(); b=b.double()
    if scheme=='exact': return F.conv2d(a,w,b,padding=k//2)
    a_hi=bf16(a); a_lo=bf16(a-a_hi); w_hi=bf16(w); w_lo=bf16(w-w_hi)
    out=F.conv2d(a_hi,w_hi,None,padding=k//2)+F.conv2d(a_hi,w_lo,None,padding=k//2)
    if scheme=='bf16x3': out=out+F.conv2d(a_lo,w_hi,None,padding=k//2)
    elif scheme=='2pass': pass
    elif scheme.startswith('fp8'):
        adt={'e5m2':torch.float8_e5m2,'e4m3':torch.float8_e4m3fn}[scheme.split('_')[1]]
        sa=2.0**9
        ws=2.0**np.floor(np.log2(224.0/w.abs().max().item()))   # static per-layer scale into e4m3 range
        aq=q8(a_lo*sa, adt)/sa
        wq=q8(w*ws, torch.float8_e4m3fn)/ws
        out=out+F.conv2d(aq,wq,None,padding=k//2)
    elif scheme=='bf16x1': out=F.conv2d(a_hi,w_hi,None,padding=k//2)
    if scheme.startswith('fp8x2'):
        # both corrections in fp8: a_hi(e5m2) x w_lo(e4m3, static scale) replaces the bf16 a_hi x w_lo pass
        adt={'e5m2':torch.float8_e5m2,'e4m3':torch.float8_e4m3fn}[scheme.split('_')[1]]
        out=F.conv2d(a_hi,w_hi,None,padding=k//2)
        sa=2.0**9
        ws=2.0**np.floor(np.log2(224.0/w.abs().max().item()))
        wls=2.0**np.floor(np.log2(224.0/max(w_lo.abs().max().item(),1e-30)))
        out=out+F.conv2d(q8(a_lo*sa,adt)/sa, q8(w*ws,torch.float8_e4m3fn)/ws,None,padding=k//2)
        out=out+F.conv2d(q8(a_hi,adt), q8(w_lo*wls,torch.float8_e4m3fn)/wls,None,padding=k//2)
        out=out+b.view(1,-1,1,1)
    return out if scheme.startswith('fp8x2') else out+b.view(1,-1,1,1)
def forward(sd,x,wb,he,gc,scheme):
    t=torch.cat([x,wb,he,gc],1).double()
    for name,_,_,k in ofw.CMG_LAYERS[:-1]:
        t=F.relu(conv_scheme(t,sd[f'cmg.{name}.weight'],sd[f'cmg.{name}.bias'],k,scheme))
    cm=torch.sigmoid(conv_scheme(t,sd['cmg.conv8.weight'],sd['cmg.conv8.bias'],3,scheme))
    out=0
    for r,(ref,o) in enumerate(zip(ofw.REFINERS,(wb,he,gc))):
        u=torch.cat([x,o],1).double()
        for name,_,_,k in ofw.REFINER_LAYERS:
            u=F.relu(conv_scheme(u,sd[f'{ref}.{name}.weight'],sd[f'{ref}.{name}.bias'],k,scheme))
        out=out+u*cm[:,r:r+1]
    return out
for gain in (1.0,3.0):
  for seed in (0,1):
    sd=ofw.synthetic_state_dict(seed,gain)
    rgb=ofw.synthetic_image(seed+10,96,96,'smooth' if seed==0 else 'noise')
    wbi,gci,hei=opre.transform(rgb)
    ten=lambda a: torch.from_numpy(a.astype(np.float32)/255).permute(2,0,1)[None]
    ins=[ten(rgb),ten(wbi),ten(hei),ten(gci)]
    ref=forward(sd,*ins,'exact')
    res={}
    for sch in ('bf16x3','fp8_e5m2','fp8x2_e5m2','fp8x2_e4m3'):
        o=forward(sd,*ins,sch)
        res[sch]=((o-ref).abs().max()/ref.abs().max()).item()
    print('gain',gain,'seed',seed,' '.join(f'{k}={v:.2e}' for k,v in res.items()),flush=True)
