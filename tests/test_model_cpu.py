"""CPU: host logic of the product (module wiring, views/strides, hand-written backward passes) against the oracle,
with the kernels replaced by tests/emu_backend.py (same bf16-operand / fp32-accumulate precision plan)."""
import pytest
import torch

from oracle import mirror_oracle as O
import emu_backend
import parity


@pytest.fixture(autouse=True)
def _emu():
    emu_backend.use()
    yield
    emu_backend.release()


# name: (cfg, B, seed, loss tolerance).  The north-star tolerances are loss rel <= 1e-3, cosine >= 0.999, grad rel-L2 <= 1e-2
# at the reference's width (E=768, 128-d latent, 3000 prototypes).  The two miniature configs (8..16-d latents, 24..40
# prototypes, 3..5 samples) average far fewer rounding errors in the quadratic style / cluster terms, so their loss
# tolerance is 2e-3; embedding and gradient tolerances are the north-star ones everywhere.
CASES = {
    "small_e192": (dict(Dw=64, Dr=100, E=192, N=150, style_hidden=64, style_out=48, latent=16, prototypes=40), 3, 11, 2e-3),
    "ragged_e192": (dict(Dw=40, Dr=77, E=192, N=97, style_hidden=32, style_out=24, latent=8, prototypes=24), 5, 12, 2e-3),
    "e768_n300": (dict(Dw=96, Dr=300, E=768, N=300, prototypes=3000), 2, 13, 1e-3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_full_step_matches_oracle(name):
    over, B, seed, loss_tol = CASES[name]
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    model = parity.build_product(cfg, sd)
    p = parity.run_product(model, wsi, rna, noise)
    o = parity.run_oracle(sd, wsi, rna, noise)
    r = parity.compare(p, o)
    assert r["mask_equal"]
    assert r["loss_rel"]["total"] <= loss_tol, r["loss_rel"]
    assert min(r["cos"].values()) >= 0.999, r["cos"]                # embedding cosine >= 0.999
    assert r["grad_rel_l2"] <= 1e-2, r["grad_rel_l2"]               # gradient rel-L2 <= 1e-2


def test_state_dict_keys_match_oracle_contract():
    cfg = O.default_cfg(Dw=64, Dr=100, E=192, N=150, prototypes=40)
    sd = O.make_state_dict(cfg, 0)
    model = parity.build_product(cfg, sd)
    got = model.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert tuple(got[k].shape) == tuple(sd[k].shape), k


def test_factories_accept_timm_injected_kwargs(caplog):
    """train_mirror.py:689-694 builds the model through timm's create_model, which injects pretrained / pretrained_cfg /
    pretrained_cfg_overlay (and passes checkpoint_path / scriptable separately): the factories filter unknown kwargs with the same
    warning as the reference (models/mirror.py:1045-1053) instead of failing"""
    import logging
    from mirror_b200.models import mirror, mirror_dual_encoder
    with caplog.at_level(logging.WARNING):
        m = mirror(wsi_embed_dim=16, rna_embed_dim=20, embed_dim=24, wsi_num_tokens=9, num_prototypes=5, pretrained=False,
                   pretrained_cfg=None, pretrained_cfg_overlay=None)
    assert "Filtered model kwargs" in caplog.text and "pretrained" in caplog.text
    assert m.wsi_encoder.retention_gene_embed.shape == (1, 10, 24) and m.prototypes.weight.shape == (5, 24)
    d = mirror_dual_encoder(wsi_embed_dim=16, rna_embed_dim=20, embed_dim=24, pretrained=False)
    assert hasattr(d, "wsi_encoder") and hasattr(d, "rna_encoder")
