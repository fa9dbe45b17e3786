"""Variable-length bags: MIRROR.forward_varlen against the oracle = the reference algorithm at B = 1 per slide
with wsi_num_tokens = N_i (oracle.mirror_forward_varlen).  CPU: emulated kernels; GPU: the CUDA path at E = 768."""
import pytest
import torch

from oracle import mirror_oracle as O
import emu_backend
import parity


def _problem(over, lengths, seed):
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    g = torch.Generator().manual_seed(seed + 1)
    B = len(lengths)
    bags = [torch.randn(n, cfg["Dw"], generator=g) for n in lengths]
    rna = torch.randn(B, cfg["Dr"], generator=g)
    noise = {"wsi_mask": [torch.rand(1, n, generator=g) for n in lengths],
             "rna_mask": torch.rand(B, cfg["E"], generator=g),
             "wsi_eps": torch.randn(B, cfg["latent"], generator=g), "rna_eps": torch.randn(B, cfg["latent"], generator=g)}
    return cfg, sd, bags, rna, noise


def _run(cfg, sd, bags, rna, noise, device):
    from mirror_b200.losses import MIRRORLoss
    model = parity.build_product(cfg, sd, device).eval()
    dev = lambda t: t.to(device)
    nz = {k: ([dev(x) for x in v] if isinstance(v, list) else dev(v)) for k, v in noise.items()}
    out = model.forward_varlen([dev(b) for b in bags], dev(rna), 0.75, 0.75, noise=nz)
    losses = MIRRORLoss()(*out)
    losses[0].backward()
    grads = {n: (p.grad.detach().float().cpu() if p.grad is not None else torch.zeros_like(p).cpu()) for n, p in model.named_parameters()}
    p = ([o.detach().float().cpu() for o in out], [l.detach().float().cpu() for l in losses], grads)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    oo = O.mirror_forward_varlen(sdo, bags, rna, noise)
    ol = O.mirror_loss(oo)
    og = O.grads_of(ol[0], sdo)
    o = ([x.detach().float() for x in oo], [l.detach().float() for l in ol], {k: v.float() for k, v in og.items()})
    return parity.compare(p, o), p, o


def _check(r, loss_tol):
    assert r["mask_equal"]
    assert r["loss_rel"]["total"] <= loss_tol, r["loss_rel"]
    assert min(r["cos"].values()) >= 0.999, r["cos"]
    assert r["grad_rel_l2"] <= 1e-2, r["grad_rel_l2"]


def test_varlen_matches_per_slide_reference_cpu():
    emu_backend.use()
    try:
        # lengths: a repeated one (runs as a batch of 2 with PER-SLIDE pinv scales), a perfect square, the table maximum
        cfg, sd, bags, rna, noise = _problem(dict(Dw=40, Dr=77, E=192, N=97, style_hidden=32, style_out=24, latent=8, prototypes=24),
                                             [60, 97, 60, 36], 51)
        r, p, o = _run(cfg, sd, bags, rna, noise, "cpu")
        _check(r, 2e-3)  # miniature configuration (see test_model_cpu.py)
        assert p[0][1].shape == (1, 60 + 97 + 60 + 36, 192) and p[0][3].shape == (1, 253)
        assert int(p[0][3].sum()) == sum(n - int(n * 0.25) for n in (60, 97, 60, 36))
    finally:
        emu_backend.release()


def test_varlen_rejects_bags_longer_than_the_position_table():
    emu_backend.use()
    try:
        cfg, sd, bags, rna, noise = _problem(dict(Dw=40, Dr=77, E=192, N=50, style_hidden=32, style_out=24, latent=8, prototypes=24), [60, 40], 52)
        model = parity.build_product(cfg, sd).eval()
        with pytest.raises(ValueError):
            model.forward_varlen(bags, rna)
    finally:
        emu_backend.release()


@pytest.mark.gpu
def test_varlen_matches_per_slide_reference_gpu():
    cfg, sd, bags, rna, noise = _problem(dict(Dw=96, Dr=300, E=768, N=1500, prototypes=3000), [1500, 700, 1024, 700, 333], 53)
    r, p, o = _run(cfg, sd, bags, rna, noise, "cuda")
    _check(r, 1e-3)
    # equal-length slides share launches but not the pinv scale: the fixed-length path (batch-global scale) differs measurably,
    # the varlen path equals the per-slide reference
    assert p[0][0].shape == (5, 768)
