"""Shared parity harness: build the product model from an oracle state_dict, run product and oracle on identical
inputs / noise, and report the north-star metrics (loss rel. error, embedding cosine, gradient rel. L2)."""
import torch

from oracle import mirror_oracle as O


def build_product(cfg, sd, device="cpu"):
    from mirror_b200.models import MIRROR
    model = MIRROR(wsi_embed_dim=cfg["Dw"], rna_embed_dim=cfg["Dr"], embed_dim=cfg["E"], wsi_num_tokens=cfg["N"],
                   rna_mlp_ratio=cfg["mlp_ratio"], rna_norm_layer="layernorm", rna_act_layer="gelu",
                   wsi_retention_decoder_depth=cfg["wsi_dec_depth"], rna_encoder_depth=cfg["rna_depth"],
                   rna_retention_decoder_depth=cfg["rna_dec_depth"], style_mlp_hidden_dim=cfg["style_hidden"],
                   style_mlp_out_dim=cfg["style_out"], style_latent_dim=cfg["latent"], num_prototypes=cfg["prototypes"])
    missing, unexpected = model.load_state_dict(sd, strict=True)
    return model.to(device)


def run_product(model, wsi, rna, noise, train=False, weights=None):
    from mirror_b200.losses import MIRRORLoss
    model.train(train)
    model.zero_grad(set_to_none=True)
    out = model(wsi, rna, 0.75, 0.75, noise=noise)
    loss_fn = MIRRORLoss() if weights is None else MIRRORLoss(True, *weights)
    losses = loss_fn(*out)
    losses[0].backward()
    grads = {n: (p.grad.detach().float().cpu() if p.grad is not None else torch.zeros_like(p).cpu()) for n, p in model.named_parameters()}
    return [o.detach().float().cpu() for o in out], [l.detach().float().cpu() for l in losses], grads


def run_oracle(sd, wsi, rna, noise, dtype=torch.float32, drop=None):
    sdo = {k: v.detach().cpu().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = O.mirror_forward(sdo, wsi.cpu().to(dtype), rna.cpu().to(dtype), {k: v.cpu().to(dtype) for k, v in noise.items()}, drop=drop)
    losses = O.mirror_loss(out)
    grads = O.grads_of(losses[0], sdo)
    return [o.detach().float() for o in out], [l.detach().float() for l in losses], {k: v.float() for k, v in grads.items()}


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def min_cos(a, b):
    a, b = a.double().flatten(0, -2) if a.dim() > 1 else a.double()[None], b.double().flatten(0, -2) if b.dim() > 1 else b.double()[None]
    return float(torch.nn.functional.cosine_similarity(a, b, dim=-1).min())


OUT_NAMES = ["wsi_align", "wsi_ret", "wsi_ret_target", "wsi_mask", "wsi_score", "wsi_mu", "wsi_logstd",
             "rna_align", "rna_ret", "rna_ret_target", "rna_mask", "rna_score", "rna_mu", "rna_logstd", "logit_scale"]
LOSS_NAMES = ["total", "align", "wsi_ret", "rna_ret", "style", "cluster"]


def compare(p, o, verbose=True):
    """p, o = (outs, losses, grads).  Returns dict of north-star metrics."""
    p_out, p_loss, p_g = p
    o_out, o_loss, o_g = o
    res = {"loss_rel": {n: abs(float(a) - float(b)) / (abs(float(b)) + 1e-30) for n, a, b in zip(LOSS_NAMES, p_loss, o_loss)}}
    res["cos"] = {n: min_cos(a, b) for n, a, b in zip(OUT_NAMES, p_out, o_out) if a.dim() >= 2 and n not in ("wsi_mask", "rna_mask")}
    res["out_rel"] = {n: rel(a, b) for n, a, b in zip(OUT_NAMES, p_out, o_out)}
    res["mask_equal"] = bool(torch.equal(p_out[3], o_out[3]) and torch.equal(p_out[10], o_out[10]))
    keys = sorted(o_g)
    assert set(p_g) == set(o_g), set(p_g) ^ set(o_g)
    gp = torch.cat([p_g[k].flatten() for k in keys])
    go = torch.cat([o_g[k].flatten() for k in keys])
    res["grad_rel_l2"] = rel(gp, go)
    res["grad_rel_per_param"] = {k: rel(p_g[k], o_g[k]) for k in keys}
    if verbose:
        print("loss rel:", {k: f"{v:.2e}" for k, v in res["loss_rel"].items()})
        print("min cosine:", {k: f"{v:.6f}" for k, v in res["cos"].items()})
        print("masks equal:", res["mask_equal"], " grad rel-L2 (all params): %.3e" % res["grad_rel_l2"])
        worst = sorted(res["grad_rel_per_param"].items(), key=lambda kv: -kv[1])[:8]
        print("worst per-param grad rel:", [(k, f"{v:.2e}", f"|g|={float(o_g[k].norm()):.2e}") for k, v in worst])
    return res
