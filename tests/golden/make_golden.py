"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run once, in the container that has ``/root/reference``::

    python tests/golden/make_golden.py

It imports ``/root/reference/waternet/{data,net}.py`` and ``hubconf.py`` by file
path (never copied), feeds them seeded synthetic inputs / weights
(``oracle.forward.synthetic_image`` / ``synthetic_state_dict``) and stores what
they return.  The fixtures travel to the GPU box; the reference does not.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("WATERNET_REFERENCE", "/root/reference")

from oracle.forward import synthetic_image, synthetic_state_dict  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    pkg = types.ModuleType("waternet")
    pkg.__path__ = [os.path.join(REF, "waternet")]
    sys.modules["waternet"] = pkg
    data = _load("waternet.data", os.path.join(REF, "waternet", "data.py"))
    net = _load("waternet.net", os.path.join(REF, "waternet", "net.py"))
    hub = _load("ref_hubconf", os.path.join(REF, "hubconf.py"))
    return data, net, hub


PRE_CASES = [
    ("noise_112x112", 0, 112, 112, "noise"),
    ("smooth_112x112", 1, 112, 112, "smooth"),
    ("noise_113x117", 2, 113, 117, "noise"),
    ("smooth_64x96", 3, 64, 96, "smooth"),
    ("noise_115x112", 4, 115, 112, "noise"),
    ("smooth_120x200", 5, 120, 200, "smooth"),
]

# (name, [(img seed, kind)], H, W, weight seed, gain)
FWD_CASES = [
    ("c1_1x112x112", [(10, "smooth")], 112, 112, 0, 1.0),
    ("n2_112x112", [(11, "noise"), (12, "smooth")], 112, 112, 0, 1.0),
    ("gain3_1x40x56", [(13, "noise")], 40, 56, 1, 3.0),
    # BASELINE.json configs[1]: batch 16, 112x112 (stress weights; mixed noise / smooth frames)
    ("c2_16x112x112", [(100 + i, "noise" if i % 4 == 0 else "smooth") for i in range(16)], 112, 112, 2, 3.0),
]


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    data, net, hub = load_reference()
    for name, seed, h, w, kind in PRE_CASES:
        rgb = synthetic_image(seed, h, w, kind)
        wb, gc, he = data.transform(rgb)
        np.savez_compressed(os.path.join(HERE, f"preprocess_{name}.npz"), rgb=rgb, wb=wb, gc=gc, he=he)
        print("preprocess", name, rgb.shape)

    preprocess, postprocess, model = hub.waternet(pretrained=False)
    model.eval()
    for name, imgs, h, w, wseed, gain in FWD_CASES:
        sd = synthetic_state_dict(wseed, gain)
        model.load_state_dict(sd, strict=True)
        rgbs = [synthetic_image(s, h, w, k) for s, k in imgs]
        parts = [preprocess(r) for r in rgbs]
        x, wb, he, gc = (torch.cat([p[i] for p in parts], dim=0) for i in range(4))
        with torch.no_grad():
            out = model(x, wb, he, gc)
        post = postprocess(out)
        np.savez_compressed(
            os.path.join(HERE, f"forward_{name}.npz"),
            rgb=np.stack(rgbs),
            out=out.numpy(),
            post=post,
            weight_seed=np.int64(wseed),
            gain=np.float64(gain),
            in_strides=np.array(parts[0][0].stride(), dtype=np.int64),
        )
        print("forward", name, tuple(out.shape), float(out.max()))


if __name__ == "__main__":
    main()
