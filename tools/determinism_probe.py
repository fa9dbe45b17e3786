"""Run-to-run determinism of one step's gradients (same inputs, same noise): the only legitimate run-to-run differences are the
summation orders of fp32 atomics (split-K weight gradients, loss reductions), i.e. ~1e-6 relative.  Anything larger is a race.
usage: PYTHONPATH=. python tools/determinism_probe.py [repeats]"""
import sys
import torch
sys.path.insert(0, "tests")
import parity
from oracle import mirror_oracle as O

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cfg = O.default_cfg(Dw=96, Dr=300, E=768, N=300, prototypes=300)
sd = O.make_state_dict(cfg, 43)
wsi, rna = O.make_inputs(4, cfg["N"], cfg["Dw"], cfg["Dr"], 143)
noise = {k: v.cuda() for k, v in O.make_noise(4, cfg["N"], cfg["E"], cfg["latent"], 243).items()}
model = parity.build_product(cfg, sd, "cuda")
for name, sl in (("all 4", slice(0, 4)), ("slides 0-1", slice(0, 2)), ("slides 2-3", slice(2, 4))):
    ref = None
    worst = (0.0, "")
    for i in range(reps):
        out, l, g = parity.run_product(model, wsi[sl].cuda(), rna[sl].cuda(), {k: v[sl] for k, v in noise.items()})
        if ref is None:
            ref = (out, g)
            continue
        w = max(((float((g[k] - ref[1][k]).norm() / (ref[1][k].norm() + 1e-30)), k) for k in g))
        wo = max(((float((a - b).norm() / (b.norm() + 1e-30)), f"out{j}") for j, (a, b) in enumerate(zip(out, ref[0]))))
        worst = max(worst, w, wo)
        if i == reps - 1:
            top = sorted(((float((g[k] - ref[1][k]).norm() / (ref[1][k].norm() + 1e-30)), k, float(ref[1][k].norm())) for k in g), reverse=True)[:6]
            print("   last run, top grads:", [(f"{a:.1e}", k, f"|g|={n:.2e}") for a, k, n in top])
            print("   last run, outputs  :", [f"{float((a - b).norm() / (b.norm() + 1e-30)):.1e}" for a, b in zip(out, ref[0])])
    print(f"{name:12s} worst run-to-run rel diff over {reps} runs: {worst[0]:.3e}  ({worst[1]})")
