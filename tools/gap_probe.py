"""Where does the device sit idle inside a step?  Kernel records of one step (torch profiler / CUPTI) sorted by start time:
idle gap after each kernel, aggregated by the kernel that precedes the gap.  usage: PYTHONPATH=. python tools/gap_probe.py [B] [graph]"""
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

from mirror_b200.losses import MIRRORLoss
from mirror_b200.models import MIRROR

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
N, Dw, Dr = 2048, 768, 10234
dev = torch.device("cuda")
torch.manual_seed(0)
model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0, rna_norm_layer="layernorm",
               rna_act_layer="gelu").to(dev).train()
loss_fn = MIRRORLoss().to(dev)
wsi, rna = torch.randn(B, N, Dw, device=dev), torch.randn(B, Dr, device=dev)


def eager():
    for p in model.parameters():
        p.grad = None
    loss_fn(*model(wsi, rna, 0.75, 0.75))[0].backward()


step = eager
if use_graph:
    from mirror_b200.step import GraphedStep
    gs = GraphedStep(model, loss_fn, (wsi, rna))
    step = lambda: gs.step(wsi, rna)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if str(getattr(e, "device_type", "")).endswith("CUDA") and e.time_range is not None:
        ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
busy = sum(b - a for a, b, _ in ev)
span = ev[-1][1] - ev[0][0]
gaps = defaultdict(lambda: [0, 0.0])
big = []
for (a0, b0, n0), (a1, b1, n1) in zip(ev, ev[1:]):
    g = max(0.0, a1 - b0)
    key = n0.replace("void ", "").replace("mb::(anonymous namespace)::", "").split("(")[0][:60]
    gaps[key][0] += 1
    gaps[key][1] += g
    big.append((g, key, n1.replace("void ", "").replace("mb::(anonymous namespace)::", "").split("(")[0][:50]))
print(f"{len(ev)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms ({'graph' if use_graph else 'eager'})")
print("idle time after each kernel name (top 25):")
for k, (c, g) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {g / 1e3:7.3f} ms  {c:4d} x {g / c:6.1f} us   {k}")
print("largest single gaps:")
for g, k, n in sorted(big, reverse=True)[:12]:
    print(f"  {g:8.1f} us after {k}  -> before {n}")
