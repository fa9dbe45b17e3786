"""How much of the data-parallel collectives of one step is hidden behind compute?  Run with torchrun on >= 2 GPUs; rank 0 profiles one
replay of the graphed step (CUPTI kernel records through torch.profiler: nsys is not installed in this image) and reports, for every
NCCL kernel, its duration and the fraction of it during which one of this repo's kernels was running on another stream.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_overlap_probe.py [slides_per_gpu] [global_negatives 0|1]"""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mirror_b200.losses import MIRRORLoss  # noqa: E402
from mirror_b200.models import MIRROR  # noqa: E402
from mirror_b200.step import GraphedStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
glob = len(sys.argv) > 2 and sys.argv[2] == "1"
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
N, Dw, Dr = 2048, 768, 10234
model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0, rna_norm_layer="layernorm",
               rna_act_layer="gelu").to(dev).train()
wsi, rna = torch.randn(B, N, Dw, device=dev), torch.randn(B, Dr, device=dev)
gs = GraphedStep(model, MIRRORLoss(global_negatives=glob).to(dev), (wsi, rna), group=dist.group.WORLD)
for _ in range(3):
    gs.step(wsi, rna)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs.step(wsi, rna)
    torch.cuda.synchronize()
if rank == 0:
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
          if str(getattr(e, "device_type", "")).endswith("CUDA") and e.time_range is not None]
    ev.sort()
    nccl = [e for e in ev if "nccl" in e[2].lower()]
    comp = [e for e in ev if "nccl" not in e[2].lower() and "memcpy" not in e[2].lower() and "memset" not in e[2].lower()]
    span = ev[-1][1] - ev[0][0]
    print(f"# {dist.get_world_size()} GPUs x {B} slides, global negatives {'on' if glob else 'off'}: one replay of the graphed step on rank 0, "
          f"{len(ev)} device activities over {span / 1e3:.2f} ms; {len(nccl)} NCCL kernels")
    print(f"# {'NCCL kernel':60s} {'start ms':>9s} {'dur us':>8s} {'hidden behind compute':>22s}")
    tot = hid = 0.0
    for a, b, n in nccl:
        o = sum(max(0.0, min(b, d) - max(a, c)) for c, d, _ in comp if d > a and c < b)
        o = min(o, b - a)
        tot += b - a
        hid += o
        print(f"  {n[:60]:60s} {(a - ev[0][0]) / 1e3:9.2f} {b - a:8.1f} {100 * o / max(b - a, 1e-9):21.0f}%")
    print(f"# NCCL kernel time {tot / 1e3:.2f} ms per step, {100 * hid / max(tot, 1e-9):.0f} % of it concurrent with compute kernels; "
          f"step span {span / 1e3:.2f} ms")
gs.close()
dist.barrier()
dist.destroy_process_group()
