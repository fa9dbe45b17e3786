"""Print the north-star parity table (GPU product vs CPU oracle) for the test configurations; run under gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import parity  # noqa: E402
from oracle import mirror_oracle as O  # noqa: E402
from test_model_gpu import CASES  # noqa: E402

print("| case | B | loss rel (total) | worst term rel | min cosine | grad rel-L2 | worst param grad rel |")
print("|---|---|---|---|---|---|---|")
for name, (over, B, seed, tol) in CASES.items():
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    model = parity.build_product(cfg, sd, "cuda")
    p = parity.run_product(model, wsi.cuda(), rna.cuda(), {k: v.cuda() for k, v in noise.items()})
    o = parity.run_oracle(sd, wsi, rna, noise)
    r = parity.compare(p, o, verbose=False)
    big = {k: v for k, v in r["grad_rel_per_param"].items() if float(o[2][k].norm()) > 1e-3 * max(float(g.norm()) for g in o[2].values())}
    wk = max(big, key=big.get)
    print(f"| {name} | {B} | {r['loss_rel']['total']:.1e} | {max(r['loss_rel'].values()):.1e} | {min(r['cos'].values()):.6f} | "
          f"{r['grad_rel_l2']:.2e} | {big[wk]:.1e} ({wk}) |", flush=True)
