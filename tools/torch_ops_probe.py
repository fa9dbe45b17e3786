"""List the torch (non-product) kernels of one training step with shapes and python call sites:  python tools/torch_ops_probe.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from mirror_b200.losses import MIRRORLoss  # noqa: E402
from mirror_b200.models import MIRROR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, Dw, Dr = (int(sys.argv[2]) if len(sys.argv) > 2 else 2048), 768, 10234
dev = torch.device("cuda")
torch.manual_seed(0)
model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0, rna_norm_layer="layernorm",
               rna_act_layer="gelu").to(dev).train()
loss_fn = MIRRORLoss().to(dev)
wsi, rna = torch.randn(B, N, Dw, device=dev), torch.randn(B, Dr, device=dev)


def step():
    for p in model.parameters():
        p.grad = None
    losses = loss_fn(*model(wsi, rna, 0.75, 0.75))
    losses[0].backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    self_t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if e.key.startswith("aten::") and self_t > 20:
        stack = [s for s in e.stack if "mirror_b200" in s or "bench" in s or "losses" in s] or list(e.stack)[:4]
        rows.append((self_t, e.count, e.key, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack[:3])))
rows.sort(reverse=True)
print(f"{'self us':>9s} {'n':>4s}  op / shapes / python site")
for r in rows[:45]:
    print(f"{r[0]:9.0f} {r[1]:4d}  {r[2]:28s} {r[3]}\n{'':16s}{r[4]}")

# ---- device-side view of the same step: time per kernel name (CUDA-event free, from the profiler's kernel records)
from collections import defaultdict  # noqa: E402

agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if getattr(ev, "device_type", None) is not None and str(ev.device_type).endswith("CUDA") and ev.name and not ev.name.startswith("aten::"):
        dt = getattr(ev, "device_time", 0) or getattr(ev, "cuda_time", 0)
        name = ev.name.replace("void ", "").replace("mb::(anonymous namespace)::", "").split("(")[0]
        agg[name][0] += 1
        agg[name][1] += dt
tot = sum(v[1] for v in agg.values())
print(f"\n{len(agg)} kernel names, {tot / 1e3:.2f} ms of device time in the step")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:48]:
    print(f"{t / 1e3:8.3f} ms {c:5d}  {name[:110]}")

# ---- host side: how long does the CPU need to enqueue one step (no synchronisation inside the loop)?
import time  # noqa: E402

torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"\nhost enqueue time per step: {(t1 - t0) / 5 * 1e3:.1f} ms; device drained {(t2 - t1) * 1e3:.1f} ms after the last enqueue; wall per step {(t2 - t0) / 5 * 1e3:.1f} ms")
