"""The training step as one unit (mirror_b200/step.py): flat parameters, fused Adam / clip / clamp tail,
CUDA-graph capture.  CPU: the step sequence against torch's own optimizer on the emulated kernels.  GPU: the Adam kernel
against torch.optim.Adam, graph replay against the eager launch sequence, fresh dropout masks on every replay."""
import copy
import math

import pytest
import torch

from oracle import mirror_oracle as O
import emu_backend
import parity

CFG = dict(Dw=40, Dr=77, E=192, N=60, style_hidden=32, style_out=24, latent=8, prototypes=24)


def _problem(B=3, seed=21):
    cfg = O.default_cfg(**CFG)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 100)
    noise = O.make_noise(B, cfg["N"], cfg["E"], cfg["latent"], seed + 200)
    return cfg, sd, wsi, rna, noise


def _trainer_steps(model, wsi, rna, noise, n, lr, clip):
    """the reference trainer's step sequence (train_mirror.py:1133-1136, 1221-1230, 1254-1256) with torch's optimizer"""
    from mirror_b200.losses import MIRRORLoss
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    loss_fn = MIRRORLoss()
    losses = []
    for _ in range(n):
        with torch.no_grad():
            model.prototypes.weight.data = torch.nn.functional.normalize(model.prototypes.weight.data, p=2, dim=1)
        opt.zero_grad(set_to_none=True)
        ls = loss_fn(*model(wsi, rna, 0.75, 0.75, noise=noise))
        ls[0].backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        opt.step()
        with torch.no_grad():
            model.logit_scale.clamp_(0, math.log(100))
        losses.append(float(ls[0].detach()))
    return losses


def test_step_sequence_matches_trainer_with_torch_adam_cpu():
    from mirror_b200.losses import MIRRORLoss
    from mirror_b200.step import GraphedStep
    emu_backend.use()
    try:
        cfg, sd, wsi, rna, noise = _problem()
        sd["logit_scale"] = torch.tensor(math.log(100.0) - 1e-4)  # the clamp must bite
        ref = parity.build_product(cfg, sd).eval()
        ours = parity.build_product(cfg, sd).eval()
        gs = GraphedStep(ours, MIRRORLoss(), (wsi, rna), optimizer=dict(lr=1e-3), clip_grad=0.5, noise=noise, graph=False)
        # one step: identical gradients go into both optimizers -> the updates agree to rounding
        want = _trainer_steps(ref, wsi, rna, noise, 1, 1e-3, 0.5)
        got = [float(gs.step(wsi, rna))]
        st = gs.read_stats()
        assert set(st) == set(GraphedStep.STATS) and st["grad_norm"] > 0
        assert abs(got[0] - want[0]) <= 1e-6 * abs(want[0])
        assert float(ours.logit_scale) <= math.log(100.0) + 1e-7
        for (n1, p1), (n2, p2) in zip(ours.named_parameters(), ref.named_parameters()):
            assert n1 == n2
            assert float((p1.detach() - p2.detach()).abs().max()) <= 2e-6, n1   # lr = 1e-3: a wrong update would be ~1e-3
        # further steps: the bf16 operand rounding of the (emulated) kernels turns 1e-8 parameter differences into 1e-4 loss
        # differences, so the trajectories are compared loosely
        ref2 = parity.build_product(cfg, sd).eval()
        want = _trainer_steps(ref2, wsi, rna, noise, 3, 1e-3, 0.5)
        got += [float(gs.step(wsi, rna)) for _ in range(2)]
        for a, b in zip(got, want):
            assert abs(a - b) <= 5e-3 * abs(b), (got, want)
        assert got[2] < got[0]
        # the parameters are views of ONE flat buffer and the state_dict contract is untouched
        assert all(p.data.untyped_storage().data_ptr() == gs.flat.data.untyped_storage().data_ptr() for p in ours.parameters())
        assert set(ours.state_dict()) == set(sd)
    finally:
        emu_backend.release()


@pytest.mark.gpu
@pytest.mark.parametrize("wd,decoupled,clip", [(0.0, False, False), (0.01, False, True), (0.05, True, False)])
def test_adam_kernel_matches_torch(wd, decoupled, clip):
    from mirror_b200 import kernels as K
    dev = "cuda"
    g_ = torch.Generator(device=dev).manual_seed(3)
    n = 1_000_003
    p = torch.randn(n + 1, device=dev, generator=g_)[:n].clone()
    ref = torch.nn.Parameter(p.clone())
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=3e-4, weight_decay=wd)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    lr = torch.full((1,), 3e-4, device=dev)
    t = torch.zeros(1, device=dev)
    coef = torch.ones(1, device=dev)
    for it in range(4):
        g = torch.randn(n, device=dev, generator=g_) * (10.0 if it == 1 else 0.1)
        ref.grad = g.clone()
        if clip:
            torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        ss = K.grad_sumsq(g) if clip else None
        K.tail_scalars_(ss, 1.0 if clip else 0.0, coef, t)
        K.adam_step_(p, g, m, v, lr, 0.9, 0.999, 1e-8, wd, decoupled, t, coef if clip else None)
        if clip:
            assert abs(float(ss) - float((g.double() ** 2).sum())) <= 1e-4 * float(ss)
    torch.cuda.synchronize()
    assert float(t) == 4.0
    assert float((p - ref.data).abs().max()) <= 2e-6, float((p - ref.data).abs().max())


def _gpu_model(sd, cfg):
    return parity.build_product(cfg, sd, device="cuda")


@pytest.mark.gpu
def test_graph_replay_equals_eager_step():
    from mirror_b200.losses import MIRRORLoss
    from mirror_b200.step import GraphedStep
    cfg, sd, wsi, rna, noise = _problem(B=4)
    dev = "cuda"
    wsi, rna = wsi.to(dev), rna.to(dev)
    noise = {k: v.to(dev) for k, v in noise.items()}
    eager = _gpu_model(sd, cfg).eval()
    _, e_loss, e_g = parity.run_product(eager, wsi, rna, noise)
    model = _gpu_model(sd, cfg).eval()
    gs = GraphedStep(model, MIRRORLoss(), (wsi, rna), noise=noise)
    assert gs.graph is not None and gs.kernels_per_replay > 100
    for _ in range(2):  # replays are idempotent without an optimizer
        loss = gs.step(wsi, rna)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(e_loss[0])) <= 1e-5 * abs(float(e_loss[0]))
        gp = torch.cat([p.grad.flatten().cpu() for _, p in sorted(model.named_parameters())])
        go = torch.cat([e_g[k].flatten() for k in sorted(e_g)])
        # fp32 atomics reorder sums and the bf16 Moore-Penrose backward amplifies that ("Run-to-run determinism")
        assert parity.rel(gp, go) <= 5e-3, parity.rel(gp, go)
    st = gs.read_stats()
    for name, want in zip(parity.LOSS_NAMES, e_loss):
        key = {"align": "alignment", "wsi_ret": "wsi_retention", "rna_ret": "rna_retention"}.get(name, name)
        assert abs(st[key] - float(want)) <= 1e-5 * abs(float(want)) + 1e-7
    # other inputs through the same graph
    wsi2 = torch.randn_like(wsi)
    l2 = float(gs.step(wsi2, rna))
    _, e2, _ = parity.run_product(eager, wsi2, rna, noise)
    assert abs(l2 - float(e2[0])) <= 1e-5 * abs(float(e2[0]))


@pytest.mark.gpu
def test_graph_replay_draws_fresh_dropout_and_noise_and_trains():
    from mirror_b200.losses import MIRRORLoss
    from mirror_b200.step import GraphedStep
    cfg, sd, wsi, rna, _ = _problem(B=4)
    wsi, rna = wsi.cuda(), rna.cuda()
    model = _gpu_model(sd, cfg).train()
    p0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    gs = GraphedStep(model, MIRRORLoss(), (wsi, rna), optimizer=dict(lr=1e-3), clip_grad=1.0)
    for n, p in model.named_parameters():  # the warm-up steps before the capture left no trace
        if n != "prototypes.weight":
            assert torch.equal(p.detach(), p0[n]), n
    losses = [float(gs.step(wsi, rna)) for _ in range(6)]
    torch.cuda.synchronize()
    assert len(set(losses)) == len(losses)             # fresh masks / noise / dropout (and moving weights) on every replay
    assert int(gs.epoch) >= 6 and float(gs.t) == 6.0
    assert all(math.isfinite(x) for x in losses)
    moved = sum(float((p.detach() - p0[n]).abs().max()) > 0 for n, p in model.named_parameters())
    assert moved == len(p0)
    w = model.prototypes.weight.detach()
    # dropout-only variation: no optimizer, fixed noise -> the loss still changes between replays in train mode
    model2 = _gpu_model(sd, cfg).train()
    noise = {k: v.cuda() for k, v in O.make_noise(4, cfg["N"], cfg["E"], cfg["latent"], 5).items()}
    gs2 = GraphedStep(model2, MIRRORLoss(), (wsi, rna), noise=noise)
    a, b = float(gs2.step(wsi, rna)), float(gs2.step(wsi, rna))
    assert a != b
    model2.eval()
