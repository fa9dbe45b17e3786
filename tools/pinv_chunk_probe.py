"""Does running the Moore-Penrose chain chunk-major (a few dozen (slide, head) matrices through all six iterations before the
next chunk) keep the iterates in the 126 MB L2 and lift the 384^3 products off the HBM bound?  Device time of the forward
chain (24 products) for several chunk sizes, each sequence replayed from a CUDA graph (no host time in the figure).
usage: PYTHONPATH=. python tools/pinv_chunk_probe.py [BH]"""
import sys
import torch
from mirror_b200 import kernels as K

BH = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = 384
dev = "cuda"
BF16 = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
a2 = torch.softmax(torch.randn(BH, m, m, device=dev, generator=g), -1)
a2_16 = a2.to(BF16)
z0 = (a2.transpose(1, 2) / (a2.abs().sum(-1).max() * a2.abs().sum(-2).max())).to(BF16).contiguous()
T = lambda x: x.transpose(-1, -2)
bufs = [[torch.empty(BH, m, m, device=dev, dtype=BF16) for _ in range(4)] for _ in range(6)]


def chain(chunk):
    for c0 in range(0, BH, chunk):
        s = slice(c0, min(BH, c0 + chunk))
        z = z0[s]
        x = a2_16[s]
        for it in range(6):
            Em, G1, Fm, zn = (b[s] for b in bufs[it])
            K.gemm(x, T(z), out_bf16=Em, alpha=-1.0, diag=1.0)
            K.gemm(Em, T(Em), out_bf16=G1, alpha=0.25, res=Em)
            K.gemm(Em, T(G1), out_bf16=Fm, res=Em)
            K.gemm(z, T(Fm), out_bf16=zn, res=z)
            z = zn


flop = 24 * 2.0 * BH * m ** 3
for chunk in [int(c) for c in __import__("os").environ.get("CHUNKS", "512,256,148,128,99,74,64,49,37,25").split(",")]:
    if chunk > BH:
        continue
    chain(chunk)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain(chunk)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunk {chunk:4d}: {ms:7.3f} ms  {flop / ms / 1e9:7.0f} TFLOP/s   ({-(-BH // chunk) * 24} launches, {chunk * 6} tiles per launch)")
    ref = bufs[5][3].float().clone() if chunk == BH else ref
    if chunk != BH:
        print("         max diff vs unchunked:", float((bufs[5][3].float() - ref).abs().max()))
