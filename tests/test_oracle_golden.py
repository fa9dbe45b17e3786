"""CPU: the oracle restatement reproduces the golden vectors that
oracle/pin_against_reference.py recorded from the REAL reference sources."""
import os

import numpy as np
import pytest
import torch

from oracle import mirror_oracle as O
from oracle.pin_against_reference import CASES, run_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sub(a, n=60000):
    a = a.reshape(-1) if a.size > n else a
    return a[:: max(1, a.size // n)][:n] if a.size > n else a


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    over, B, seed = CASES[name]
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    out, losses, grads = run_oracle(cfg, sd, wsi, rna, noise)
    gold = np.load(os.path.join(GOLDEN, f"mirror_{name}.npz"))
    np.testing.assert_allclose(torch.stack(losses).numpy(), gold["losses"], rtol=2e-5)
    for i, o in enumerate(out):
        np.testing.assert_allclose(_sub(o.numpy()), gold[f"out{i}"], rtol=2e-4, atol=2e-5)
    gn = np.array([float(grads[k].norm()) for k in sorted(grads)])
    np.testing.assert_allclose(gn, gold["grad_norms"], rtol=1e-3, atol=1e-7)
    gs = np.concatenate([grads[k].flatten()[:8].numpy() for k in sorted(grads)])
    np.testing.assert_allclose(gs, gold["grad_samples"], rtol=5e-3, atol=1e-6)


def test_contrastive_golden():
    gold = np.load(os.path.join(GOLDEN, "contrastive.npz"))
    g = torch.Generator().manual_seed(5)
    q, k = torch.randn(37, 64, generator=g), torch.randn(37, 64, generator=g)
    for sym in (False, True):
        qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
        l = O.info_nce(qq, kk, 0.1, sym)
        l.backward()
        np.testing.assert_allclose(float(l), gold[f"nce_sym{int(sym)}"][0], rtol=1e-6)
        np.testing.assert_allclose(qq.grad.numpy(), gold[f"nce_sym{int(sym)}_dq"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(kk.grad.numpy(), gold[f"nce_sym{int(sym)}_dk"], rtol=1e-4, atol=1e-7)
    qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    s = torch.tensor(1 / 0.07, requires_grad=True)
    l = O.clip_loss(qq, kk, s)
    l.backward()
    np.testing.assert_allclose(float(l), gold["clip"][0], rtol=1e-6)
    np.testing.assert_allclose(qq.grad.numpy(), gold["clip_dq"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(s.grad), gold["clip_ds"][0], rtol=1e-4)


def test_symmetric_infonce_equals_cliploss_on_normalised_inputs():
    # one fused kernel serves both losses
    g = torch.Generator().manual_seed(9)
    q, k = torch.randn(19, 32, generator=g), torch.randn(19, 32, generator=g)
    a = O.info_nce(q, k, 0.07, True)
    b = O.clip_loss(torch.nn.functional.normalize(q, dim=-1), torch.nn.functional.normalize(k, dim=-1), torch.tensor(1 / 0.07))
    assert abs(float(a) - float(b)) < 1e-6


def test_masking_keeps_expected_count():
    x = torch.randn(4, 50, 8)
    noise = torch.rand(4, 50)
    y, m = O.random_masking(x, torch.zeros(1, 8), 0.75, noise)
    assert int(m.sum()) == 4 * (50 - int(50 * 0.25))
    assert torch.equal(y[m == 0], x[m == 0]) and float(y[m == 1].abs().sum()) == 0.0
