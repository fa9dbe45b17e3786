#!/usr/bin/env python
"""The inner loop of the reference's train_mirror.py (lines 1125-1280) on mirror_b200's step-level API, synthetic data.

    python examples/pretrain_loop.py [--steps 20] [--batch 16] [--patches 2048]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/pretrain_loop.py --global-negatives

What replaces what:
  dataset __getitem__ resampling + collate + .to(device)   ->  data.resample_indices + data.SlidePrefetcher (pinned ring, copy
  (datasets/dataset_pretrain.py:150-167, :1138-1139)            stream, device-side gather over the packed bf16 patch features)
  prototype normalise, autocast forward, loss, backward,    ->  step.GraphedStep: one CUDA graph per step, fused Adam, gradient
  optimizer.step(), logit_scale clamp, seven .item() calls      all-reduce overlapped with the backward, ONE device->host copy
  (train_mirror.py:1133-1136, 1144-1230, 1254-1274)             of the loss scalars
The unchanged trainer keeps working too (mirror_b200.models.mirror + mirror_b200.losses.MIRRORLoss are drop-ins); this loop is
what a maintainer would switch to for the small-batch / multi-GPU regime, where ~630 kernel launches per step from Python
(34 ms of host time) are the bottleneck.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mirror_b200.data import SlidePrefetcher, resample_indices  # noqa: E402
from mirror_b200.losses import MIRRORLoss  # noqa: E402
from mirror_b200.models import mirror  # noqa: E402
from mirror_b200.step import GraphedStep  # noqa: E402


def synthetic_loader(steps, batch, patches, wsi_dim, rna_dim, seed):
    """what a cached TCGAWSIRNAPretrainDataset + DataLoader would hand over: per slide a [M_i, Dw] bf16 feature matrix with its own
    number of patches (packed along dim 0), the resampling indices of this epoch, and the fp32 RNA vector"""
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        lengths = rng.randint(patches // 2, patches * 2, size=batch)  # fewer or more patches than wsi_num_tokens
        packed = torch.randn(int(lengths.sum()), wsi_dim, generator=g).to(torch.bfloat16)
        index = resample_indices(lengths, patches, rng)               # np.random.choice(M_i, N, replace=M_i < N) per slide
        yield (packed, index), torch.randn(batch, rna_dim, generator=g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=16, help="slides per GPU")
    ap.add_argument("--patches", type=int, default=2048)
    ap.add_argument("--lr", type=float, default=2e-5)  # configs/pretrain/mirror.template.yaml
    ap.add_argument("--clip-grad", type=float, default=None)
    ap.add_argument("--global-negatives", action="store_true")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(0)
    wsi_dim, rna_dim = 768, 10234
    model = mirror(wsi_embed_dim=wsi_dim, rna_embed_dim=rna_dim, embed_dim=768, wsi_num_tokens=a.patches, rna_mlp_ratio=4.0,
                   rna_norm_layer="layernorm", rna_act_layer="gelu").to(dev).train()
    loss_fn = MIRRORLoss(global_negatives=a.global_negatives).to(dev)
    example = (torch.zeros(a.batch, a.patches, wsi_dim, device=dev), torch.zeros(a.batch, rna_dim, device=dev))
    step = GraphedStep(model, loss_fn, example, group=group, optimizer=dict(lr=a.lr), clip_grad=a.clip_grad)
    loader = SlidePrefetcher(synthetic_loader(a.steps, a.batch, a.patches, wsi_dim, rna_dim, 1234 + rank), dev)
    t0 = time.perf_counter()
    for i, (wsi, rna) in enumerate(loader):
        step.step(wsi, rna)
        if rank == 0 and (i % 5 == 0 or i == a.steps - 1):
            s = step.read_stats()  # one D2H copy; also the only host synchronisation of the loop
            print(f"step {i:3d}  loss {s['total']:.4f}  align {s['alignment']:.4f}  wsi_ret {s['wsi_retention']:.4f}  rna_ret {s['rna_retention']:.4f}  "
                  f"style {s['style']:.3f}  cluster {s['cluster']:.4f}  logit_scale {s['logit_scale_exp']:.3f}  |g| {s['grad_norm']:.3f}", flush=True)
    torch.cuda.synchronize()
    if rank == 0:
        dt = time.perf_counter() - t0
        print(f"{a.steps} optimizer steps, {world * a.batch * a.steps / dt:.0f} slides/s incl. Adam and the data pipeline (the synthetic features are generated on the host inside the loop) "
              f"({loader.h2d_bytes / a.steps / 1e6:.0f} MB H2D per step per GPU, {step.kernels_per_replay} kernels per graph replay)")
    step.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
