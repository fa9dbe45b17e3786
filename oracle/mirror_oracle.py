"""CPU oracle for the MIRROR pre-training step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; nothing under
``mirror_b200/`` does.  It is the *checker*, never the product.

What it is: a plain-PyTorch (fp32 or fp64, CPU or any device) functional
restatement of the reference's hot path, driven by a reference-compatible
``state_dict`` — no ``nn.Module`` classes, no timm, no nystrom_attention.
Each function cites the reference lines it follows (paths relative to the
reference checkout).

Pinning: ``oracle/pin_against_reference.py`` executes the UNMODIFIED reference
sources from the read-only mount (through ``oracle/shims``) and checks this
restatement against them (outputs, losses and all parameter gradients, to
float round-off); it also writes ``tests/golden/*.npz``.  The reference itself
has no tests or golden vectors, and the Nyström arithmetic comes
from the un-vendored PyPI package ``nystrom_attention~=0.0.14``, so that part
is "parity unpinned upstream": it follows the published algorithm.

All randomness of the reference forward (``models/mirror.py:516,630,830-833``)
is injected through ``noise`` so both sides of a parity test see the same draws.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

WSI_HEADS = 8  # models/mirror.py:302
RNA_HEADS = 12  # models/mirror.py:161,392 (MIRROR never overrides it)
PINV_ITERS = 6  # models/mirror.py:304
RES_KERNEL = 33  # nystrom_attention default residual_conv_kernel


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _lin(sd: SD, p: str, x: Tensor) -> Tensor:
    b = sd.get(p + ".bias")
    return F.linear(x, sd[p + ".weight"], b)


def _ln(sd: SD, p: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _drop(x: Tensor, key: str, drop: Optional[Dict[str, Tensor]]) -> Tensor:
    """Inverted dropout with an injected keep-mask already scaled by 1/(1-p)."""
    if drop is None or key not in drop:
        return x
    return x * drop[key]


# --------------------------------------------------------------------------
# Nyström attention (third-party; published algorithm, see module docstring)
# --------------------------------------------------------------------------
def pinv_iter(x: Tensor, iters: int = PINV_ITERS) -> Tensor:
    ax = x.abs()
    denom = ax.sum(-1).max() * ax.sum(-2).max()  # maxima over the whole tensor
    z = x.transpose(-1, -2) / denom
    eye = torch.eye(x.shape[-1], dtype=x.dtype, device=x.device)
    for _ in range(iters):
        xz = x @ z
        z = 0.25 * z @ (13 * eye - xz @ (15 * eye - xz @ (7 * eye - xz)))
    return z


def nystrom_attention(sd: SD, p: str, x: Tensor, drop=None, dkey="") -> Tensor:
    """x: [B,S,E] already layer-normed.  Call site models/mirror.py:299-312."""
    B, S, E = x.shape
    h, m = WSI_HEADS, E // 2
    d = E // h
    pad = (m - S % m) % m
    if pad:
        x = F.pad(x, (0, 0, pad, 0))  # FRONT padding with zero rows
    n = S + pad
    qkv = F.linear(x, sd[p + ".to_qkv.weight"])
    q, k, v = (t.reshape(B, n, h, d).transpose(1, 2) for t in qkv.chunk(3, -1))
    q = q * d ** -0.5
    seg = math.ceil(S / m)
    q_l = q.reshape(B, h, n // seg, seg, d).sum(3) / seg
    k_l = k.reshape(B, h, n // seg, seg, d).sum(3) / seg
    a1 = (q @ k_l.transpose(-1, -2)).softmax(-1)
    a2 = (q_l @ k_l.transpose(-1, -2)).softmax(-1)
    a3 = (q_l @ k.transpose(-1, -2)).softmax(-1)  # zero pad keys are NOT masked
    out = (a1 @ pinv_iter(a2)) @ (a3 @ v)
    out = out + F.conv2d(v, sd[p + ".res_conv.weight"], padding=(RES_KERNEL // 2, 0), groups=h)
    out = out.transpose(1, 2).reshape(B, n, E)
    out = _lin(sd, p + ".to_out.0", out)
    out = _drop(out, dkey, drop)  # Dropout(0.1) acts on the padded length n
    return out[:, -S:]


def trans_layer(sd: SD, p: str, x: Tensor, drop=None) -> Tensor:
    """models/mirror.py:311-314 (LayerNorm eps = torch default 1e-5)."""
    return x + nystrom_attention(sd, p + ".attn", _ln(sd, p + ".norm", x, 1e-5), drop, p)


def ppeg(sd: SD, p: str, x: Tensor, H: int) -> Tensor:
    """models/mirror.py:324-331."""
    B, _, C = x.shape
    cls, feat = x[:, :1], x[:, 1:]
    f = feat.transpose(1, 2).reshape(B, C, H, H)
    y = f
    for name, ksz in (("proj", 7), ("proj1", 5), ("proj2", 3)):
        y = y + F.conv2d(f, sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], padding=ksz // 2, groups=C)
    return torch.cat([cls, y.flatten(2).transpose(1, 2)], 1)


# --------------------------------------------------------------------------
# WSI encoder + decoders (FeatureTransMILHybrid)
# --------------------------------------------------------------------------
def wsi_encoder(sd: SD, wsi: Tensor, drop=None, p: str = "wsi_encoder") -> Tuple[Tensor, int]:
    """models/mirror.py:651-679.  Returns ([B,N+1,E], add_length)."""
    h = F.relu(_lin(sd, p + "._fc1.0", wsi.to(sd[p + "._fc1.0.weight"].dtype)))
    B, N, _ = h.shape
    H = int(math.ceil(math.sqrt(N)))
    add = H * H - N
    h = torch.cat([h, h[:, :add]], 1)  # wrap-around square padding
    h = torch.cat([sd[p + ".cls_token"].expand(B, -1, -1), h], 1)
    h = trans_layer(sd, p + ".layer1", h, drop)
    h = ppeg(sd, p + ".pos_layer", h, H)
    h = trans_layer(sd, p + ".layer2", h, drop)
    h = _ln(sd, p + ".norm", h, 1e-5)
    return h[:, : h.shape[1] - add], add


def cls_encoder(sd: SD, wsi: Tensor, p: str = "wsi_encoder") -> Tensor:
    """FeatureTransMIL.forward, models/mirror.py:352-380 -> cls embedding [B,E]."""
    return wsi_encoder(sd, wsi, None, p)[0][:, 0]


def random_masking(x: Tensor, mask_token: Tensor, ratio: float, noise: Tensor) -> Tuple[Tensor, Tensor]:
    """models/mirror.py:510-533 / 624-649: keep the int(N(1-r)) smallest-noise
    slots, every other slot := mask_token; mask is 1 where masked.  Equivalent
    closed form of the argsort/gather/cat/gather sequence."""
    N = x.shape[1]
    keep = int(N * (1 - ratio))
    rank = torch.argsort(torch.argsort(noise, dim=1), dim=1)  # ids_restore
    mask = (rank >= keep).to(x.dtype)
    mk = mask if x.dim() == 2 else mask.unsqueeze(-1)
    return x * (1 - mk) + mask_token.to(x.dtype) * mk, mask


def wsi_decoders(sd: SD, h: Tensor, ratio: float, noise: Tensor, drop=None, p: str = "wsi_encoder"):
    """models/mirror.py:681-706."""
    eps = 1e-6 if h.dtype == torch.float16 else 1e-12
    align = _lin(sd, p + ".alignment_head", F.normalize(h, dim=-1, eps=eps)[:, 0])
    r = _lin(sd, p + ".retention_embed", h)
    rm, mask = random_masking(r[:, 1:], sd[p + ".mask_token"][0], ratio, noise)
    r = torch.cat([r[:, :1], rm], 1) + sd[p + ".retention_gene_embed"]
    i = 0
    while f"{p}.retention_blocks.{i}.norm.weight" in sd:
        r = trans_layer(sd, f"{p}.retention_blocks.{i}", r, drop)
        i += 1
    r = _lin(sd, p + ".retention_head", _ln(sd, p + ".retention_norm", r, 1e-5))
    return align, r[:, 1:], mask


# --------------------------------------------------------------------------
# RNA encoder + decoders (TransFormerHybrid)
# --------------------------------------------------------------------------
def rna_attention(sd: SD, p: str, x: Tensor, drop=None) -> Tensor:
    """models/mirror.py:77-102: attention over the 12 chunks of ONE vector."""
    B, E = x.shape
    hd = E // RNA_HEADS
    qkv = _lin(sd, p + ".qkv", x).reshape(B, 3, RNA_HEADS, hd)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]  # each [B,12,hd]: seq=12, dim=hd
    a = ((q * hd ** -0.5) @ k.transpose(-1, -2)).softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(B, E)  # dim-major interleave
    return _drop(_lin(sd, p + ".proj", o), p + ".proj", drop)


def rna_block(sd: SD, p: str, x: Tensor, drop=None) -> Tensor:
    """models/mirror.py:149-152 (LayerNorm eps 1e-6, exact-erf GELU)."""
    x = x + rna_attention(sd, p + ".attn", _ln(sd, p + ".norm1", x, 1e-6), drop)
    y = F.gelu(_lin(sd, p + ".mlp.fc1", _ln(sd, p + ".norm2", x, 1e-6)))
    y = _drop(y, p + ".mlp.drop1", drop)
    y = _drop(_lin(sd, p + ".mlp.fc2", y), p + ".mlp.drop2", drop)
    return x + y


def rna_encoder(sd: SD, rna: Tensor, drop=None, p: str = "rna_encoder") -> Tensor:
    """models/mirror.py:283-289 with the timm Mlp embedding (:217-224)."""
    x = F.gelu(_lin(sd, p + ".embedding.fc1", rna))
    x = _lin(sd, p + ".embedding.fc2", _ln(sd, p + ".embedding.norm", x, 1e-6))
    if p + ".gene_embed" in sd:
        x = x + sd[p + ".gene_embed"]
    i = 0
    while f"{p}.blocks.{i}.norm1.weight" in sd:
        x = rna_block(sd, f"{p}.blocks.{i}", x, drop)
        i += 1
    return _ln(sd, p + ".norm", x, 1e-6)


def rna_decoders(sd: SD, x: Tensor, ratio: float, noise: Tensor, drop=None, p: str = "rna_encoder"):
    """models/mirror.py:538-561."""
    eps = 1e-6 if x.dtype == torch.float16 else 1e-12
    align = _lin(sd, p + ".alignment_head", F.normalize(x, dim=-1, eps=eps))
    r = _lin(sd, p + ".retention_embed", x)
    r, mask = random_masking(r, sd[p + ".mask_token"][0, 0], ratio, noise)
    r = r + sd[p + ".retention_gene_embed"]
    i = 0
    while f"{p}.retention_blocks.{i}.norm1.weight" in sd:
        r = rna_block(sd, f"{p}.retention_blocks.{i}", r, drop)
        i += 1
    r = _lin(sd, p + ".retention_head", _ln(sd, p + ".retention_norm", r, 1e-6))
    return align, r, mask


# --------------------------------------------------------------------------
# style / cluster heads and the full forward
# --------------------------------------------------------------------------
def style_heads(sd: SD, emb: Tensor, eps_noise: Tensor):
    """models/mirror.py:830-858 for one modality.  ``logstd`` is a log-variance."""
    x = _lin(sd, "style_encoder_mlp.fc2", F.gelu(_lin(sd, "style_encoder_mlp.fc1", emb)))
    mu, logstd = _lin(sd, "style_mu", x), _lin(sd, "style_logstd", x)
    z = mu + torch.exp(0.5 * logstd) * eps_noise.to(mu.dtype)  # Normal.rsample
    score = F.linear(_lin(sd, "style_decoder", z), sd["prototypes.weight"])
    return score, mu, logstd


def mirror_forward(sd: SD, wsi: Tensor, rna: Tensor, noise: Dict[str, Tensor],
                   wsi_mask_ratio: float = 0.75, rna_mask_ratio: float = 0.75, drop=None):
    """MIRROR.forward, models/mirror.py:860-915 -> the 15-tuple in reference order."""
    wsi_emb, _ = wsi_encoder(sd, wsi, drop)
    wa, wr, wm = wsi_decoders(sd, wsi_emb, wsi_mask_ratio, noise["wsi_mask"], drop)
    rna_emb = rna_encoder(sd, rna.to(wsi_emb.dtype), drop)
    ra, rr, rm = rna_decoders(sd, rna_emb, rna_mask_ratio, noise["rna_mask"], drop)
    ws, wmu, wls = style_heads(sd, wsi_emb[:, 0], noise["wsi_eps"])
    rs, rmu, rls = style_heads(sd, rna_emb, noise["rna_eps"])
    return (wa, wr, wsi_emb[:, 1:], wm, ws, wmu, wls,
            ra, rr, rna_emb, rm, rs, rmu, rls, sd["logit_scale"].exp())


def mirror_forward_varlen(sd: SD, bags, rna: Tensor, noise: Dict[str, object],
                          wsi_mask_ratio: float = 0.75, rna_mask_ratio: float = 0.75):
    """Variable-length bags (no reference counterpart in batched form): every slide i with its own N_i patches
    is encoded and decoded by the reference algorithm AT B = 1 with wsi_num_tokens = N_i (own H_i, S_i, landmark group size,
    pinv scale; position table retention_gene_embed sliced to its first N_i + 1 rows), the batch-level parts (RNA encoder,
    style / cluster heads) run on the whole batch.  Ragged retention outputs are packed along the token axis as
    [1, sum N_i, E] with mask [1, sum N_i], for which losses/mirror_loss.py:98-103 is exactly the token-weighted masked MSE.
    noise["wsi_mask"] is a list of [1, N_i] tensors."""
    pfx = "wsi_encoder"
    was, wrs, wts, wms, cls = [], [], [], [], []
    for i, bag in enumerate(bags):
        emb, _ = wsi_encoder(sd, bag[None])
        sdi = dict(sd)
        sdi[pfx + ".retention_gene_embed"] = sd[pfx + ".retention_gene_embed"][:, : bag.shape[0] + 1]
        wa, wr, wm = wsi_decoders(sdi, emb, wsi_mask_ratio, noise["wsi_mask"][i])
        was.append(wa), wrs.append(wr), wts.append(emb[:, 1:]), wms.append(wm), cls.append(emb[:, 0])
    rna_emb = rna_encoder(sd, rna.to(was[0].dtype))
    ra, rr, rm = rna_decoders(sd, rna_emb, rna_mask_ratio, noise["rna_mask"])
    ws, wmu, wls = style_heads(sd, torch.cat(cls), noise["wsi_eps"])
    rs, rmu, rls = style_heads(sd, rna_emb, noise["rna_eps"])
    return (torch.cat(was), torch.cat(wrs, 1), torch.cat(wts, 1), torch.cat(wms, 1), ws, wmu, wls,
            ra, rr, rna_emb, rm, rs, rmu, rls, sd["logit_scale"].exp())


def dual_encoder_forward(sd: SD, wsi: Tensor, rna: Tensor):
    """2-output model train_pretrain.py:1119-1122 expects: FeatureTransMIL cls
    embedding (models/mirror.py:352-380) and TransFormer output (:283-289)."""
    return cls_encoder(sd, wsi), rna_encoder(sd, rna)


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
def clip_loss(w: Tensor, r: Tensor, scale: Tensor) -> Tensor:
    """losses/mirror_loss.py:37-52 (no L2 normalisation here)."""
    li = scale * w @ r.T
    lt = scale * r @ w.T
    y = torch.arange(w.shape[0], device=w.device)
    return (F.cross_entropy(li, y) + F.cross_entropy(lt, y)) / 2


def mirror_loss(out, weights=(0.5, 0.1, 0.1, 0.1, 0.2)):
    """losses/mirror_loss.py:74-135 -> (total, align, wsi_ret, rna_ret, style, cluster)."""
    wa, wr, wt, wm, ws, wmu, wls, ra, rr, rt, rm, rs, rmu, rls, scale = out
    align = clip_loss(wa, ra, scale)
    wsi_ret = (((wr - wt) ** 2).mean(-1) * wm).sum() / wm.sum()
    rna_ret = (((rr - rt) ** 2) * rm).sum() / rm.sum()
    style = 0.5 * ((wls.exp() + wmu ** 2 - 1 - wls).sum(1).mean()
                   + (rls.exp() + rmu ** 2 - 1 - rls).sum(1).mean())
    lw, lr = F.log_softmax(ws, -1), F.log_softmax(rs, -1)
    kl = lambda la, lb: (lb.exp() * (lb - la)).sum() / la.shape[0]  # kl_div(la, b, batchmean)
    cluster = 0.5 * (kl(lw, lr) + kl(lr, lw))
    a, b, c, d, e = weights
    total = a * align + b * wsi_ret + c * rna_ret + d * style + e * cluster
    return total, align, wsi_ret, rna_ret, style, cluster


def info_nce(q: Tensor, k: Tensor, temperature: float = 0.1, symmetric: bool = False,
             reduction: str = "mean") -> Tensor:
    """losses/info_nce.py:123-164, implicit-negatives branch."""
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    logits = q @ k.T
    y = torch.arange(len(q), device=q.device)
    if symmetric:
        return 0.5 * F.cross_entropy(logits / temperature, y, reduction=reduction) + \
            0.5 * F.cross_entropy(logits.T / temperature, y, reduction=reduction)
    return F.cross_entropy(logits / temperature, y, reduction=reduction)


# --------------------------------------------------------------------------
# deterministic synthetic problems (shared by tests, smoke and bench)
# --------------------------------------------------------------------------
def make_noise(B: int, N: int, E: int, latent: int, seed: int, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return {"wsi_mask": torch.rand(B, N, generator=g).to(dtype),
            "rna_mask": torch.rand(B, E, generator=g).to(dtype),
            "wsi_eps": torch.randn(B, latent, generator=g).to(dtype),
            "rna_eps": torch.randn(B, latent, generator=g).to(dtype)}


def make_inputs(B: int, N: int, Dw: int, Dr: int, seed: int, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, N, Dw, generator=g).to(dtype), torch.randn(B, Dr, generator=g).to(dtype)


def default_cfg(**over):
    cfg = dict(Dw=768, Dr=10234, E=768, N=2048, rna_depth=2, mlp_ratio=4.0, wsi_dec_depth=1,
               rna_dec_depth=1, style_hidden=512, style_out=256, latent=128, prototypes=3000)
    cfg.update(over)
    return cfg


def make_state_dict(cfg, seed: int = 0, dtype=torch.float32, qk_gain: float = 1.0) -> SD:
    """Deterministic random weights with the reference's state_dict keys and
    shapes.  Scales follow the reference's init conventions
    (xavier-uniform linears, N(0,0.02) tokens) except that biases and LayerNorm
    affine terms are made non-trivial so that every term is exercised.
    ``pin_against_reference.py`` loads the result into the real reference model
    with ``strict=True``, which pins the key/shape contract."""
    g = torch.Generator().manual_seed(seed)
    E, Dw, Dr, N = cfg["E"], cfg["Dw"], cfg["Dr"], cfg["N"]
    sd: SD = {}

    def lin(p, o, i, bias=True, gain=1.0):
        a = gain * math.sqrt(6.0 / (i + o))
        sd[p + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * a
        if bias:
            sd[p + ".bias"] = torch.randn(o, generator=g) * 0.02

    def ln(p, dim):
        sd[p + ".weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        sd[p + ".bias"] = 0.05 * torch.randn(dim, generator=g)

    def nys(p):
        ln(p + ".norm", E)
        lin(p + ".attn.to_qkv", 3 * E, E, bias=False, gain=qk_gain)
        lin(p + ".attn.to_out.0", E, E)
        sd[p + ".attn.res_conv.weight"] = (torch.rand(WSI_HEADS, 1, RES_KERNEL, 1, generator=g) * 2 - 1) / math.sqrt(RES_KERNEL)

    def block(p, hidden):
        ln(p + ".norm1", E)
        lin(p + ".attn.qkv", 3 * E, E, gain=qk_gain)
        lin(p + ".attn.proj", E, E)
        ln(p + ".norm2", E)
        lin(p + ".mlp.fc1", hidden, E)
        lin(p + ".mlp.fc2", E, hidden)

    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    w = "wsi_encoder"
    sd[w + ".cls_token"] = 0.02 * torch.randn(1, 1, E, generator=g)
    sd[w + ".mask_token"] = 0.02 * torch.randn(1, 1, E, generator=g)
    sd[w + ".retention_gene_embed"] = 0.02 * torch.randn(1, N + 1, E, generator=g)
    for name, k in (("proj", 7), ("proj1", 5), ("proj2", 3)):
        sd[f"{w}.pos_layer.{name}.weight"] = (torch.rand(E, 1, k, k, generator=g) * 2 - 1) / k
        sd[f"{w}.pos_layer.{name}.bias"] = 0.02 * torch.randn(E, generator=g)
    lin(w + "._fc1.0", E, Dw)
    nys(w + ".layer1")
    nys(w + ".layer2")
    ln(w + ".norm", E)
    lin(w + ".alignment_head", E, E)
    lin(w + ".retention_embed", E, E)
    for i in range(cfg["wsi_dec_depth"]):
        nys(f"{w}.retention_blocks.{i}")
    ln(w + ".retention_norm", E)
    lin(w + ".retention_head", E, E)

    r = "rna_encoder"
    hidden = int(E * cfg["mlp_ratio"])
    sd[r + ".gene_embed"] = 0.02 * torch.randn(1, E, generator=g)
    sd[r + ".mask_token"] = 0.02 * torch.randn(1, 1, generator=g)
    sd[r + ".retention_gene_embed"] = 0.02 * torch.randn(1, E, generator=g)
    lin(r + ".embedding.fc1", 2 * E, Dr)
    ln(r + ".embedding.norm", 2 * E)
    lin(r + ".embedding.fc2", E, 2 * E)
    for i in range(cfg["rna_depth"]):
        block(f"{r}.blocks.{i}", hidden)
    ln(r + ".norm", E)
    lin(r + ".alignment_head", E, E)
    lin(r + ".retention_embed", E, E)
    for i in range(cfg["rna_dec_depth"]):
        block(f"{r}.retention_blocks.{i}", hidden)
    ln(r + ".retention_norm", E)
    lin(r + ".retention_head", E, E)

    lin("style_encoder_mlp.fc1", cfg["style_hidden"], E)
    lin("style_encoder_mlp.fc2", cfg["style_out"], cfg["style_hidden"])
    lin("style_mu", cfg["latent"], cfg["style_out"])
    lin("style_logstd", cfg["latent"], cfg["style_out"])
    lin("style_decoder", E, cfg["latent"])
    sd["prototypes.weight"] = F.normalize(torch.randn(cfg["prototypes"], E, generator=g), dim=1)  # train_mirror.py:1133-1136
    return {k: v.to(dtype) for k, v in sd.items()}


def grads_of(total: Tensor, sd: SD) -> SD:
    names = [k for k, v in sd.items() if v.requires_grad]
    gs = torch.autograd.grad(total, [sd[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, gs)}
