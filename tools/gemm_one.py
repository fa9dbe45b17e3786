"""Run one GEMM shape a few times (for ncu):  python tools/gemm_one.py BATCH M N K {NT|NN|TN|TT} {b16|f32}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mirror_b200 import kernels as K  # noqa: E402

Bt, M, N, Kd = (int(x) for x in sys.argv[1:5])
lay, out = sys.argv[5], sys.argv[6]
g = torch.Generator(device="cuda").manual_seed(0)
bf = lambda *s: torch.randn(*s, device="cuda", generator=g).to(torch.bfloat16)
a = bf(Bt, Kd, M).transpose(-1, -2) if lay[0] == "T" else bf(Bt, M, Kd)
b = bf(Bt, Kd, N).transpose(-1, -2) if lay[1] == "N" else bf(Bt, N, Kd)
o16 = torch.empty(Bt, M, N, device="cuda", dtype=torch.bfloat16) if out == "b16" else None
o32 = torch.empty(Bt, M, N, device="cuda") if out == "f32" else None
for _ in range(6):
    K.gemm(a, b, out_bf16=o16, out_f32=o32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    K.gemm(a, b, out_bf16=o16, out_f32=o32)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{Bt}x{M}x{N}x{Kd} {lay} {out}: {ms:.3f} ms {2 * Bt * M * N * Kd / ms / 1e9:.0f} TFLOP/s")
