"""Train mode: the fused hashed dropout is reproduced exactly by injecting the same keep-masks into the oracle.
CPU (emulated kernels) always; the same check on the GPU under `-m gpu`."""
import math

import pytest
import torch

import emu_backend
import parity
from mirror_b200 import ops
from oracle import mirror_oracle as O


def _masks(cfg, B, seeds):
    """keep-masks in the order the product draws its seeds (mirror_b200/models/mirror.py forward order)."""
    E, N = cfg["E"], cfg["N"]
    H = int(math.ceil(math.sqrt(N)))
    S_enc, S_dec = H * H + 1, N + 1
    m, hid = E // 2, int(E * cfg["mlp_ratio"])
    it = iter(seeds)
    drop = {}

    def nys(key, S):
        n = S + (m - S % m) % m
        full = torch.zeros(B, n, E)
        full[:, n - S:] = emu_backend.keep_mask(next(it), (B, S, E), 0.1)
        drop[key] = full

    nys("wsi_encoder.layer1", S_enc)
    nys("wsi_encoder.layer2", S_enc)
    nys("wsi_encoder.retention_blocks.0", S_dec)
    for blk in ("rna_encoder.blocks.0", "rna_encoder.blocks.1", "rna_encoder.retention_blocks.0"):
        drop[blk + ".attn.proj"] = emu_backend.keep_mask(next(it), (B, E), 0.1)
        drop[blk + ".mlp.drop1"] = emu_backend.keep_mask(next(it), (B, hid), 0.1)
        drop[blk + ".mlp.drop2"] = emu_backend.keep_mask(next(it), (B, E), 0.1)
    return drop


def _run(device):
    over, B, seed = dict(Dw=64, Dr=100, E=192, N=150, style_hidden=64, style_out=48, latent=16, prototypes=40), 4, 41
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    model = parity.build_product(cfg, sd, device)
    ops._seed_state.update(base=1234, n=0)
    seeds = []
    ops._seed_state.update(base=1234, n=0)
    for _ in range(12):
        seeds.append(ops.next_seed())
    ops._seed_state.update(base=1234, n=0)
    p = parity.run_product(model, wsi.to(device), rna.to(device), {k: v.to(device) for k, v in noise.items()}, train=True)
    o = parity.run_oracle(sd, wsi, rna, noise, drop=_masks(cfg, B, seeds))
    r = parity.compare(p, o)
    assert r["loss_rel"]["total"] <= 2e-3 and min(r["cos"].values()) >= 0.999 and r["grad_rel_l2"] <= 1e-2
    # and dropout really happened: eval-mode loss differs
    e = parity.run_product(model, wsi.to(device), rna.to(device), {k: v.to(device) for k, v in noise.items()}, train=False)
    assert abs(float(e[1][0]) - float(p[1][0])) > 1e-4


def test_train_mode_dropout_matches_oracle_cpu():
    emu_backend.use()
    try:
        _run("cpu")
    finally:
        emu_backend.release()


@pytest.mark.gpu
def test_train_mode_dropout_matches_oracle_gpu():
    _run("cuda")


def test_hash_dropout_rate():
    m = emu_backend.keep_mask(7, (200000,), 0.1)
    assert abs(float((m == 0).float().mean()) - 0.1) < 3e-3
