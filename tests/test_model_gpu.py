"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, against
the golden vectors recorded from the real reference, and size-independent properties at the benchmark shapes.
North-star tolerances (BASELINE.json): loss rel <= 1e-3, embedding cosine >= 0.999, gradient rel-L2 <= 1e-2."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import mirror_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "small_e192": (dict(Dw=64, Dr=100, E=192, N=150, style_hidden=64, style_out=48, latent=16, prototypes=40), 3, 11, 2e-3),
    "ragged_e192": (dict(Dw=40, Dr=77, E=192, N=97, style_hidden=32, style_out=24, latent=8, prototypes=24), 5, 12, 2e-3),
    "e768_n300": (dict(Dw=96, Dr=300, E=768, N=300, prototypes=3000), 2, 13, 1e-3),
    # C1 of SURVEY.md §8: the reference's own CPU-runnable configuration
    "c1_b4_n2048": (dict(Dw=1024, Dr=10234, E=768, N=2048), 4, 14, 1e-3),
}


def _problem(over, B, seed):
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    return cfg, sd, wsi, rna, noise


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_full_step_matches_oracle(name):
    over, B, seed, loss_tol = CASES[name]
    cfg, sd, wsi, rna, noise = _problem(over, B, seed)
    model = parity.build_product(cfg, sd, "cuda")
    p = parity.run_product(model, wsi.cuda(), rna.cuda(), _cuda(noise))
    o = parity.run_oracle(sd, wsi, rna, noise)
    r = parity.compare(p, o)
    assert r["mask_equal"]
    assert r["loss_rel"]["total"] <= loss_tol, r["loss_rel"]
    assert min(r["cos"].values()) >= 0.999, r["cos"]
    assert r["grad_rel_l2"] <= 1e-2, r["grad_rel_l2"]


@pytest.mark.parametrize("name", ["small_e192", "e768_n300"])
def test_against_reference_golden(name):
    """golden vectors written by oracle/pin_against_reference.py from the UNMODIFIED reference sources"""
    from oracle.pin_against_reference import CASES as PIN
    over, B, seed = PIN[name]
    cfg, sd, wsi, rna, noise = _problem(over, B, seed)
    model = parity.build_product(cfg, sd, "cuda")
    out, losses, grads = parity.run_product(model, wsi.cuda(), rna.cuda(), _cuda(noise))
    gold = np.load(os.path.join(GOLDEN, f"mirror_{name}.npz"))
    gl = gold["losses"]
    assert abs(float(losses[0]) - gl[0]) / abs(gl[0]) <= 2e-3
    for i, o in enumerate(out):
        a = o.numpy()
        a = a if a.size <= 60000 else a.reshape(-1)[:: max(1, a.size // 60000)][:60000]
        g = gold[f"out{i}"].reshape(-1)
        cos = float(np.dot(a.reshape(-1), g) / (np.linalg.norm(a) * np.linalg.norm(g) + 1e-30))
        assert cos >= 0.999, (i, cos)
    gn = np.array([float(grads[k].norm()) for k in sorted(grads)])
    # per-parameter norms: 2 % relative plus a floor of 1e-3 of the largest norm (a 192-element bias gradient summed over
    # B = 3 contrastive rows sits at the 2 % edge in bf16; the north-star bound -- global rel-L2 <= 1e-2 -- is checked above)
    G = gold["grad_norms"]
    np.testing.assert_allclose(gn, G, rtol=2e-2, atol=1e-3 * G.max())


def test_contrastive_losses_against_reference_golden():
    from mirror_b200.losses import ClipLoss, InfoNCE
    gold = np.load(os.path.join(GOLDEN, "contrastive.npz"))
    g = torch.Generator().manual_seed(5)
    q, k = torch.randn(37, 64, generator=g), torch.randn(37, 64, generator=g)
    for sym in (False, True):
        qq, kk = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
        l = InfoNCE(temperature=0.1, symmetric=sym)(qq, kk)
        l.backward()
        assert abs(float(l) - gold[f"nce_sym{int(sym)}"][0]) <= 1e-4 * abs(gold[f"nce_sym{int(sym)}"][0])
        np.testing.assert_allclose(qq.grad.cpu().numpy(), gold[f"nce_sym{int(sym)}_dq"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(kk.grad.cpu().numpy(), gold[f"nce_sym{int(sym)}_dk"], rtol=2e-3, atol=2e-5)
    qq, kk = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    s = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
    l = ClipLoss()(qq, kk, s)
    l.backward()
    assert abs(float(l) - gold["clip"][0]) <= 1e-4 * abs(gold["clip"][0])
    np.testing.assert_allclose(qq.grad.cpu().numpy(), gold["clip_dq"], rtol=2e-3, atol=1e-4)
    assert abs(float(s.grad) - gold["clip_ds"][0]) <= 2e-3 * abs(gold["clip_ds"][0])


def test_infonce_validation_errors_match_reference():
    from mirror_b200.losses import InfoNCE
    f = InfoNCE()
    q = torch.zeros(4, 8, device="cuda")
    with pytest.raises(ValueError):
        f(q[0], q)
    with pytest.raises(ValueError):
        f(q, q[:3])
    with pytest.raises(ValueError):
        f(q, torch.zeros(4, 9, device="cuda"))


@pytest.mark.parametrize("B", [256, 2048])
def test_contrastive_large_batch_properties(B):
    """C5 sizes: symmetric InfoNCE == ClipLoss on normalised inputs; gradient rows of dq sum to zero against k-mean direction"""
    from mirror_b200.losses import ClipLoss, InfoNCE
    g = torch.Generator().manual_seed(B)
    q, k = torch.randn(B, 512, generator=g).cuda(), torch.randn(B, 512, generator=g).cuda()
    a = InfoNCE(temperature=0.07, symmetric=True)(q, k)
    qn, kn = torch.nn.functional.normalize(q, dim=-1), torch.nn.functional.normalize(k, dim=-1)
    b = ClipLoss()(qn, kn, torch.tensor(1 / 0.07, device="cuda"))
    assert abs(float(a) - float(b)) <= 1e-4 * abs(float(b))
    ref = O.info_nce(q.cpu().double(), k.cpu().double(), 0.07, True)
    assert abs(float(a) - float(ref)) <= 1e-3 * abs(float(ref))


def test_dual_encoder_matches_oracle():
    from mirror_b200.models import MIRRORDualEncoder
    cfg, sd, wsi, rna, _ = _problem(dict(Dw=96, Dr=300, E=768, N=300), 3, 21)
    model = MIRRORDualEncoder(cfg["Dw"], cfg["Dr"], cfg["E"], rna_mlp_ratio=cfg["mlp_ratio"], rna_norm_layer="layernorm", rna_act_layer="gelu")
    own = model.state_dict()
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)
    model = model.cuda().eval()
    w, r = model(wsi.cuda(), rna.cuda())
    ow, orr = O.dual_encoder_forward(sd, wsi, rna)
    assert parity.min_cos(w.detach().cpu(), ow) >= 0.999 and parity.min_cos(r.detach().cpu(), orr) >= 0.999


def test_benchmark_shape_properties():
    """C3 shape (N=2048, Dw=768, E=768, Dr=10234) at B=8: properties that do not need the oracle."""
    cfg, sd, wsi, rna, noise = _problem(dict(Dw=768, Dr=10234, E=768, N=2048), 8, 31)
    model = parity.build_product(cfg, sd, "cuda")
    out, losses, grads = parity.run_product(model, wsi.cuda(), rna.cuda(), _cuda(noise))
    assert all(torch.isfinite(o).all() for o in out) and all(torch.isfinite(l) for l in losses)
    assert all(torch.isfinite(g).all() for g in grads.values())
    keep = int(2048 * 0.25)
    assert int(out[3].sum()) == 8 * (2048 - keep) and int(out[10].sum()) == 8 * (768 - int(768 * 0.25))
    w = (0.5, 0.1, 0.1, 0.1, 0.2)
    assert abs(float(losses[0]) - sum(wi * float(li) for wi, li in zip(w, losses[1:]))) <= 1e-5 * abs(float(losses[0]))
    # eval-mode forward is repeatable up to the summation order of the fp32 atomics (split-K, loss reductions)
    out2, losses2, _ = parity.run_product(model, wsi.cuda(), rna.cuda(), _cuda(noise))
    assert abs(float(losses2[0]) - float(losses[0])) <= 1e-6 * abs(float(losses[0]))
    # slides are independent in eval mode apart from the batch-global pinv scale: ~1e-4 at random init
    outh, _, _ = parity.run_product(model, wsi[:4].cuda(), rna[:4].cuda(), {k: v[:4].cuda() for k, v in noise.items()})
    assert parity.rel(outh[2], out[2][:4]) <= 5e-3
