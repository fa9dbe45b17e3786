"""Pin oracle/mirror_oracle.py against the real reference and write golden files.

Run HERE (the container that has /root/reference mounted):
    python oracle/pin_against_reference.py            # check + (re)write tests/golden/*.npz
It (1) builds the reference ``MIRROR`` (models/mirror.py:720) through
oracle/shims, (2) loads ``make_state_dict`` weights with strict=True (pins the
state_dict contract), (3) runs reference forward + ``MIRRORLoss`` + backward and
the oracle restatement on identical inputs/noise, asserting agreement to float
round-off, and (4) stores the reference's outputs as golden vectors.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mirror_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (cfg overrides, B, seed)  -- "small" is kernel-compatible (d=24, m=96)
    "small_e192": (dict(Dw=64, Dr=100, E=192, N=150, style_hidden=64, style_out=48, latent=16, prototypes=40), 3, 11),
    "ragged_e192": (dict(Dw=40, Dr=77, E=192, N=97, style_hidden=32, style_out=24, latent=8, prototypes=24), 2, 12),
    "e768_n300": (dict(Dw=96, Dr=300, E=768, N=300, prototypes=3000), 2, 13),
}


def run_reference(mods, cfg, sd, wsi, rna, noise, ratios=(0.75, 0.75)):
    mm, ml, _ = mods
    model = mm.mirror(wsi_embed_dim=cfg["Dw"], rna_embed_dim=cfg["Dr"], embed_dim=cfg["E"],
                      wsi_num_tokens=cfg["N"], rna_mlp_ratio=cfg["mlp_ratio"],
                      rna_norm_layer="layernorm", rna_act_layer="gelu",
                      style_mlp_hidden_dim=cfg["style_hidden"], style_mlp_out_dim=cfg["style_out"],
                      style_latent_dim=cfg["latent"], num_prototypes=cfg["prototypes"]).eval()
    model.load_state_dict(sd, strict=True)
    reset, _ = ref_loader.pin_noise(mm, model, noise)
    reset()
    out = model(wsi, rna, *ratios)
    losses = ml.MIRRORLoss()(*out)
    losses[0].backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    return [o.detach() for o in out], [l.detach() for l in losses], grads


def run_oracle(cfg, sd, wsi, rna, noise, dtype=torch.float32, ratios=(0.75, 0.75)):
    sdo = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = O.mirror_forward(sdo, wsi.to(dtype), rna.to(dtype), {k: v.to(dtype) for k, v in noise.items()}, *ratios)
    losses = O.mirror_loss(out)
    grads = O.grads_of(losses[0], sdo)
    return [o.detach() for o in out], [l.detach() for l in losses], grads


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    mods = ref_loader.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    worst = 0.0
    for name, (over, B, seed) in CASES.items():
        cfg = O.default_cfg(**over)
        sd = O.make_state_dict(cfg, seed)
        wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
        noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
        r_out, r_loss, r_g = run_reference(mods, cfg, sd, wsi, rna, noise)
        o_out, o_loss, o_g = run_oracle(cfg, sd, wsi, rna, noise)
        errs = [rel(a, b) for a, b in zip(o_out, r_out)] + [rel(a, b) for a, b in zip(o_loss, r_loss)]
        gerr = {k: rel(o_g[k], r_g[k]) for k in r_g}
        assert set(r_g) == set(sd), "state_dict key mismatch"
        gw = max(gerr.values())
        print(f"{name}: max out/loss rel {max(errs):.2e}; max grad rel {gw:.2e} ({max(gerr, key=gerr.get)})")
        worst = max(worst, max(errs), gw)
        assert max(errs) < 2e-5 and gw < 5e-4, "oracle restatement disagrees with the reference"
        # fp64 oracle vs fp32 reference: shows the fp32 round-off floor of the reference itself
        d_out, d_loss, d_g = run_oracle(cfg, sd, wsi, rna, noise, torch.float64)
        print(f"   fp64-oracle vs fp32-reference: loss rel {rel(d_loss[0], r_loss[0]):.2e}, "
              f"grad rel-L2 {rel(torch.cat([d_g[k].flatten() for k in sorted(r_g)]), torch.cat([r_g[k].flatten() for k in sorted(r_g)])):.2e}")
        gold = {"losses": torch.stack(r_loss).numpy()}
        for i, o in enumerate(r_out):
            a = o.numpy()
            gold[f"out{i}"] = a if a.size <= 60000 else a.reshape(-1)[:: max(1, a.size // 60000)][:60000].copy()
        gold["grad_norms"] = np.array([float(r_g[k].norm()) for k in sorted(r_g)], dtype=np.float64)
        gold["grad_samples"] = np.concatenate([r_g[k].flatten()[:8].numpy() for k in sorted(r_g)])
        np.savez_compressed(os.path.join(GOLDEN, f"mirror_{name}.npz"), **gold)

    # InfoNCE / ClipLoss golden (losses import torch only — run the real files)
    _, ml, mi = mods
    g = torch.Generator().manual_seed(5)
    q, k = torch.randn(37, 64, generator=g), torch.randn(37, 64, generator=g)
    gold = {}
    for sym in (False, True):
        qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
        l = mi.InfoNCE(temperature=0.1, symmetric=sym)(qq, kk)
        l.backward()
        lo = O.info_nce(q, k, 0.1, sym)
        assert rel(lo, l.detach()) < 1e-6
        gold[f"nce_sym{int(sym)}"] = np.array([float(l)])
        gold[f"nce_sym{int(sym)}_dq"] = qq.grad.numpy()
        gold[f"nce_sym{int(sym)}_dk"] = kk.grad.numpy()
    qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    s = torch.tensor(1 / 0.07, requires_grad=True)
    l = ml.ClipLoss()(qq, kk, s)
    l.backward()
    assert rel(O.clip_loss(q, k, s.detach()), l.detach()) < 1e-6
    gold["clip"] = np.array([float(l)]); gold["clip_dq"] = qq.grad.numpy(); gold["clip_dk"] = kk.grad.numpy()
    gold["clip_ds"] = np.array([float(s.grad)])
    np.savez_compressed(os.path.join(GOLDEN, "contrastive.npz"), **gold)
    print("pinned; worst rel err", worst)


if __name__ == "__main__":
    main()
