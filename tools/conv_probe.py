"""Time the res_conv kernels at the benchmark shape:  python tools/conv_probe.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mirror_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
E, hd, n = 768, 8, 2304
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, n, 3 * E, device="cuda", generator=g).to(torch.bfloat16)
do = torch.randn(B, n, E, device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn(hd, 33, device="cuda", generator=g)
dw = torch.zeros(hd, 33, device="cuda")


def timed(name, f):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:30s} {e0.elapsed_time(e1) / 10:8.3f} ms", flush=True)


timed("res_conv_fwd", lambda: K.res_conv_fwd(qkv, w))
timed("res_conv_bwd (data + wgrad)", lambda: K.res_conv_bwd(do, qkv, w, dw))
H = 46
x = torch.randn(B, H * H + 1, E, device="cuda", generator=g)
dy = torch.randn(B, H * H + 1, E, device="cuda", generator=g)
w7, w5, w3 = (torch.randn(E, 1, k, k, device="cuda", generator=g) for k in (7, 5, 3))
b7, b5, b3 = (torch.randn(E, device="cuda", generator=g) for _ in range(3))
y, wm = K.ppeg_fwd(x, w7, w5, w3, b7, b5, b3, H)
gr = [torch.zeros_like(t) for t in (w7, w5, w3, b7, b5, b3)]
timed("ppeg_fwd (merge + stencil)", lambda: K.ppeg_fwd(x, w7, w5, w3, b7, b5, b3, H))
timed("ppeg_bwd (data + wgrad)", lambda: K.ppeg_bwd(dy, x, wm, H, *gr))
