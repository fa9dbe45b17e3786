"""CPU, world_size 2, gloo: the drop-in model trains under the reference's own data-parallel mechanism
(torch DistributedDataParallel, train_mirror.py:811-813).  Kernels are emulated (tests/emu_backend.py); what is
checked is the host logic: hand-written autograd functions + DDP's bucketed gradient averaging give exactly the mean of
the per-rank gradients, with rank-local contrastive negatives as in the reference."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(Dw=40, Dr=77, E=192, N=60, style_hidden=32, style_out=24, latent=8, prototypes=24)


def _problem(rank):
    from oracle import mirror_oracle as O
    cfg = O.default_cfg(**CFG)
    sd = O.make_state_dict(cfg, 5)
    wsi, rna = O.make_inputs(2, cfg["N"], cfg["Dw"], cfg["Dr"], 300 + rank)
    noise = O.make_noise(2, cfg["N"], cfg["E"], cfg["latent"], 400 + rank)
    return cfg, sd, wsi, rna, noise


def _local_grads(rank):
    import parity
    cfg, sd, wsi, rna, noise = _problem(rank)
    model = parity.build_product(cfg, sd)
    _, losses, grads = parity.run_product(model, wsi, rna, noise)
    return losses, grads


def _worker(rank, world, port, q):
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import emu_backend
    import parity
    from mirror_b200.losses import MIRRORLoss
    emu_backend.use()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, sd, wsi, rna, noise = _problem(rank)
    model = parity.build_product(cfg, sd).eval()
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    out = ddp(wsi, rna, 0.75, 0.75, noise=noise)
    MIRRORLoss()(*out)[0].backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    if rank == 0:
        q.put({k: v.numpy() for k, v in grads.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_two_ranks_average_local_gradients():
    import emu_backend
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    emu_backend.use()
    try:
        g0, g1 = _local_grads(0)[1], _local_grads(1)[1]
    finally:
        emu_backend.release()
    for k in g0:
        want = 0.5 * (g0[k] + g1[k])
        err = float((torch.from_numpy(got[k]) - want).norm() / (want.norm() + 1e-12))
        assert err < 1e-5, (k, err)
