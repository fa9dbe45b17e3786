"""Bisect a run-to-run difference of the backward pass: gradients w.r.t. the intermediate activations of the WSI branch are
recorded by hooks in repeated identical runs and compared (measurement only).  usage: PYTHONPATH=. python tools/determinism_bisect.py"""
import sys
import torch
sys.path.insert(0, "tests")
import parity
from oracle import mirror_oracle as O
from mirror_b200.losses import MIRRORLoss

cfg = O.default_cfg(Dw=96, Dr=300, E=768, N=300, prototypes=300)
sd = O.make_state_dict(cfg, 43)
wsi, rna = O.make_inputs(4, cfg["N"], cfg["Dw"], cfg["Dr"], 143)
noise = {k: v.cuda() for k, v in O.make_noise(4, cfg["N"], cfg["E"], cfg["latent"], 243).items()}
model = parity.build_product(cfg, sd, "cuda").eval()
enc = model.wsi_encoder
names = ["layer1", "pos_layer", "layer2", "retention_blocks.0"]
mods = {n: dict(enc.named_modules())[n] for n in names}


def run(sl):
    rec = {}
    hs = []
    for n, m in mods.items():
        hs.append(m.register_full_backward_hook(lambda mod, gin, gout, n=n: rec.__setitem__(n, (gin[0].detach().clone() if gin[0] is not None else None, gout[0].detach().clone()))))
    model.zero_grad(set_to_none=True)
    out = model(wsi[sl].cuda(), rna[sl].cuda(), 0.75, 0.75, noise={k: v[sl] for k, v in noise.items()})
    MIRRORLoss()(*out)[0].backward()
    for h in hs:
        h.remove()
    g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return rec, g


rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
for sl in (slice(0, 2), slice(2, 4)):
    ref, gref = run(sl)
    for i in range(6):
        rec, g = run(sl)
        line = []
        for n in names:
            gi, go = rec[n]
            line.append(f"{n}: d_out {rel(go, ref[n][1]):.1e} d_in {rel(gi, ref[n][0]) if gi is not None else -1:.1e}")
        worst = max((rel(g[k], gref[k]), k) for k in g)
        print(sl, "|", " | ".join(line), "| worst param", f"{worst[0]:.1e}", worst[1])
