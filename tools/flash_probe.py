"""Time the fused Nystrom softmax-product kernels at the benchmark shape (B x 8 heads, n = 2304 tokens, m = 384 landmarks,
d = 96) with CUDA events: forward K-C / K-A and the four backward launches.  Measurement only."""
import sys
import torch
from mirror_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
E, n, m, h = 768, 2304, 384, 8
d, seg = E // h, n // m
dev = "cuda"
BF16 = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(BF16)
qkv, lm, wv, rc, do, dkv, dlm16, dvc = rnd(B, n, 3 * E), rnd(B, m, 2 * E), rnd(B, h, m, d), rnd(B, n, E), rnd(B, n, E), rnd(B, h, m, d), rnd(B, m, 2 * E), rnd(B, n, E)
hv = lambda t, c0: t[:, :, c0:c0 + E].unflatten(-1, (h, d)).permute(0, 2, 1, 3)
q, k, v, ql, kl = hv(qkv, 0), hv(qkv, E), hv(qkv, 2 * E), hv(lm, 0), hv(lm, E)
alpha = d ** -0.5
o16 = torch.empty(B, n, E, device=dev, dtype=BF16)
kv = torch.empty(B, h, m, d, device=dev, dtype=BF16)
dqkv = torch.empty(B, n, 3 * E, device=dev, dtype=BF16)
dlm32 = torch.empty(B, m, 2 * E, device=dev, dtype=torch.float32)
dw = torch.empty(B, h, m, d, device=dev, dtype=BF16)
lse1 = K.flash_softmax_pv(q, kl, wv, alpha, hv(o16, 0), hv(rc, 0))
lse3 = K.flash_softmax_pv(ql, k, v, alpha, kv)
dot1 = torch.randn(B, h, n, device=dev) * 0.01
dot3 = torch.randn(B, h, m, device=dev) * 0.01
runs = {
    "fwd K-C softmax(q kl^T) W + rc": lambda: K.flash_softmax_pv(q, kl, wv, alpha, hv(o16, 0), hv(rc, 0)),
    "fwd K-A softmax(ql k^T) v": lambda: K.flash_softmax_pv(ql, k, v, alpha, kv),
    "bwd a1 rows (dq)": lambda: K.flash_bwd(q, kl, hv(do, 0), wv, alpha, lse1, dot1, False, (hv(dqkv, 0), hv(dlm16, 0), seg, 1.0 / seg)),
    "bwd a1 cols (dkl, dW)": lambda: K.flash_bwd(kl, q, wv, hv(do, 0), alpha, lse1, dot1, True, (hv(dlm32, E), None, 1, 1.0), (dw, None, 1, 1.0)),
    "bwd a3 rows (dql)": lambda: K.flash_bwd(ql, k, dkv, v, alpha, lse3, dot3, False, (hv(dlm32, 0), None, 1, 1.0)),
    "bwd a3 cols (dk, dv)": lambda: K.flash_bwd(k, ql, v, dkv, alpha, lse3, dot3, True, (hv(dqkv, E), hv(dlm16, E), seg, 1.0 / seg), (hv(dqkv, 2 * E), hv(dvc, 0), 1, 1.0)),
}
elems = B * h * n * m
only = sys.argv[2] if len(sys.argv) > 2 else ""
for name, fn in runs.items():
    if only not in name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:36s} {ms:7.3f} ms   {elems / ms / 1e6:8.1f} G softmax-elements/s")
