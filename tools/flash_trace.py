"""Event trace of CTA 0 of the flash forward kernel at the benchmark shape (measurement only).
usage: MIRROR_B200_EXTRA_NVCC_FLAGS=-DMIRROR_FLASH_TRACE PYTHONPATH=. python tools/flash_trace.py [kc|ka|a1rows|a1cols|a3rows|a3cols] [B] [skip]
(the trace hooks are compiled out of the product build)"""
import sys
import torch
from mirror_b200 import _lib, kernels as K

which = sys.argv[1] if len(sys.argv) > 1 else "kc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
E, n, m, h = 768, 2304, 384, 8
d = E // h
dev = "cuda"
BF16 = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(BF16)
qkv, lm, wv, rc = rnd(B, n, 3 * E), rnd(B, m, 2 * E), rnd(B, h, m, d), rnd(B, n, E)
hv = lambda t, c0: t[:, :, c0:c0 + E].unflatten(-1, (h, d)).permute(0, 2, 1, 3)
q, k, v, ql, kl = hv(qkv, 0), hv(qkv, E), hv(qkv, 2 * E), hv(lm, 0), hv(lm, E)
o16 = torch.empty(B, n, E, device=dev, dtype=BF16)
kv = torch.empty(B, h, m, d, device=dev, dtype=BF16)
seg = n // m
do, dkv, dlm16, dvc = rnd(B, n, E), rnd(B, h, m, d), rnd(B, m, 2 * E), rnd(B, n, E)
dqkv = torch.empty(B, n, 3 * E, device=dev, dtype=BF16)
dlm32 = torch.empty(B, m, 2 * E, device=dev, dtype=torch.float32)
dw = torch.empty(B, h, m, d, device=dev, dtype=BF16)
alpha = d ** -0.5
lse1 = K.flash_softmax_pv(q, kl, wv, alpha, hv(o16, 0), hv(rc, 0))
lse3 = K.flash_softmax_pv(ql, k, v, alpha, kv)
dot1 = torch.randn(B, h, n, device=dev) * 0.01
dot3 = torch.randn(B, h, m, device=dev) * 0.01
runs = {
    "kc": lambda: K.flash_softmax_pv(q, kl, wv, alpha, hv(o16, 0), hv(rc, 0)),
    "ka": lambda: K.flash_softmax_pv(ql, k, v, alpha, kv),
    "a1rows": lambda: K.flash_bwd(q, kl, hv(do, 0), wv, alpha, lse1, dot1, False, (hv(dqkv, 0), hv(dlm16, 0), seg, 1.0 / seg)),
    "a1cols": lambda: K.flash_bwd(kl, q, wv, hv(do, 0), alpha, lse1, dot1, True, (hv(dlm32, E), None, 1, 1.0), (dw, None, 1, 1.0)),
    "a3rows": lambda: K.flash_bwd(ql, k, dkv, v, alpha, lse3, dot3, False, (hv(dlm32, 0), None, 1, 1.0)),
    "a3cols": lambda: K.flash_bwd(k, ql, v, dkv, alpha, lse3, dot3, True, (hv(dqkv, E), hv(dlm16, E), seg, 1.0 / seg), (hv(dqkv, 2 * E), hv(dvc, 0), 1, 1.0)),
}
run = runs[which]
for _ in range(3):
    run()
cap = 1 << 16
buf = torch.zeros(cap, device=dev, dtype=torch.int64)
_lib.check(_lib.fn("mirror_debug_flash_trace")(buf.data_ptr(), cap), "trace")
run()
torch.cuda.synchronize()
_lib.check(_lib.fn("mirror_debug_flash_trace")(None, 0), "trace")
a = buf.cpu().numpy()
ev = []
for r in range(3):
    reg = a[r * (cap // 4):(r + 1) * (cap // 4)]
    ev += [(int(x) & ((1 << 48) - 1), int(x) >> 48) for x in reg[1:int(reg[0]) + 1]]
ev.sort()
cnt = len(ev)
t0 = ev[0][0]
names = {1: "P x_load", 2: "P y_load", 3: "P v_load", 10: "M tile", 14: "M s_empty ok", 11: "M S issue", 15: "M p_full ok", 12: "M PV issue", 13: "M o_full commit",
         20: "S p0 s_full", 21: "S p0 ld done", 22: "S bar1", 23: "S p1 s_full", 24: "S p1 ld done", 25: "S p_empty ok", 26: "S P stored", 27: "S bar2",
         28: "S o_full", 29: "S epi done", 40: "M tile", 41: "M SD issued", 42: "M ds_full ok", 43: "M OUT issued", 44: "M o_full commit",
         50: "S tile start", 51: "S s_full ok", 52: "S dS stored", 53: "S o_full ok", 54: "S epi done", 60: "P tile loads", 61: "P stage load", 30: "S e0 start", 31: "S epi acc in regs", 32: "S e0 stored", 33: "S tile start", 34: "S e1 ld done", 35: "S e1 stored"}
print(f"{cnt} events")
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
for t, i in ev[skip:skip + 260]:
    print(f"{t - t0:9d}  {names.get(i, i)}")
