"""Time the fused-softmax GEMM passes on the step's own shapes (head-strided q/k views):  python tools/softmax_gemm_probe.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mirror_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(os.environ.get("REPS", "10"))
E, hd, n, m = 768, 8, 2304, 384
d = E // hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, n, 3 * E, device="cuda", generator=g).to(torch.bfloat16)
lm = torch.randn(B, m, 2 * E, device="cuda", generator=g).to(torch.bfloat16)
heads = lambda t, c0: t[:, :, c0:c0 + E].unflatten(-1, (hd, d)).permute(0, 2, 1, 3)
q, kl = heads(qkv, 0), heads(lm, E)
qc, klc = q.contiguous(), kl.contiguous()


ONCE = os.environ.get("PROBE_ONCE") == "1"  # one launch per variant (ncu capture)


def timed(name, f, bytes_=0):
    if ONCE:
        f()
        return
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:42s} {ms:8.3f} ms   {bytes_ / ms / 1e9 if bytes_ else 0:7.2f} TB/s", flush=True)


ONLY = os.environ.get("PROBE_ONLY", "")
for tag, (a, b) in ({} if ONLY == "pinv" else {"strided": (q, kl)} if ONCE else {"strided": (q, kl), "contig": (qc, klc)}).items():
    o32 = torch.empty(B, hd, n, m, device="cuda")
    o16 = torch.empty(B, hd, n, m, device="cuda", dtype=torch.bfloat16)
    st = K.softmax_stats((B, hd), n, m, "cuda")
    nel = B * hd * n * m
    timed(f"s1 {tag} normal f32", lambda: K.gemm(a, b, out_f32=o32, alpha=0.1), nel * 4)
    timed(f"s1 {tag} normal bf16", lambda: K.gemm(a, b, out_bf16=o16, alpha=0.1), nel * 2)
    timed(f"s1 {tag} NULL epilogue (mode 6)", lambda: K.gemm(a, b, alpha=0.1, mode=6, stats=st))
    timed(f"s1 {tag} ROWSTATS", lambda: K.gemm(a, b, alpha=0.1, mode=K.GEMM_ROWSTATS, stats=st))
    timed(f"s1 {tag} SOFTMAX bf16", lambda: K.gemm(a, b, alpha=0.1, mode=K.GEMM_SOFTMAX, stats=st, out_bf16=o16), nel * 2)
    timed(f"s1 {tag} ROWDOT", lambda: K.gemm(a, b, mode=K.GEMM_ROWDOT, stats=st, res=o16), nel * 2)
    timed(f"s1 {tag} SOFTMAX_BWD", lambda: K.gemm(a, b, alpha=0.1, mode=K.GEMM_SOFTMAX_BWD, stats=st, res=o16, out_bf16=o16), nel * 4)
    del o32
# the Moore-Penrose products: [m, m] x [m, m], bf16 out with a bf16 residual
z = torch.randn(B, hd, m, m, device="cuda", generator=g).to(torch.bfloat16)
zo = torch.empty_like(z)
timed("pinv 384^3 NN bf16 + res", lambda: K.gemm(z, z.transpose(-1, -2), out_bf16=zo, res=z), 2 * B * hd * m * m * 2)
timed("pinv 384^3 NT bf16", lambda: K.gemm(z, z, out_bf16=zo, alpha=0.25), B * hd * m * m * 2)
timed("pinv 384^3 NN bf16 (no res)", lambda: K.gemm(z, z.transpose(-1, -2), out_bf16=zo), B * hd * m * m * 2)
timed("pinv 384^3 NT bf16 + res", lambda: K.gemm(z, z, out_bf16=zo, res=z), 2 * B * hd * m * m * 2)
timed("pinv 384^3 TN bf16", lambda: K.gemm(z.transpose(-1, -2), z.transpose(-1, -2), out_bf16=zo), B * hd * m * m * 2)
if ONLY == "av":
    # attention-value products: the n x m probabilities (906 MB) stream through once
    a1 = torch.randn(B, hd, n, m, device="cuda", generator=g).to(torch.bfloat16)
    w_ = torch.randn(B, hd, m, d, device="cuda", generator=g).to(torch.bfloat16)
    do = torch.randn(B, n, E, device="cuda", generator=g).to(torch.bfloat16)
    o16 = torch.empty(B, n, E, device="cuda", dtype=torch.bfloat16)
    doh, oh = heads(do.repeat(1, 1, 3), 0), heads(o16.repeat(1, 1, 3), 0)
    oh = o16.unflatten(-1, (hd, d)).permute(0, 2, 1, 3)
    doh = do.unflatten(-1, (hd, d)).permute(0, 2, 1, 3)
    timed("out = a1 w   [2304x96, K=384] NN", lambda: K.gemm(a1, w_.transpose(-1, -2), out_bf16=oh), B * hd * n * m * 2)
    dw = torch.empty(B, hd, m, d, device="cuda", dtype=torch.bfloat16)
    timed("dw = a1^T dO [384x96, K=2304] TN", lambda: K.gemm(a1.transpose(-1, -2), doh.transpose(-1, -2), out_bf16=dw), B * hd * n * m * 2)
    w96 = torch.randn(B, hd, d, m, device="cuda", generator=g).to(torch.bfloat16)
    stn = K.softmax_stats((B, hd), n, d, "cuda")
    timed("out = a1 w   NT NULL epilogue", lambda: K.gemm(a1, w96, mode=6, stats=stn), B * hd * n * m * 2)
    sys.exit(0)
if ONLY == "pinv":
    stz = K.softmax_stats((B, hd), m, m, "cuda")
    timed("pinv 384^3 NT NULL epilogue (mode 6)", lambda: K.gemm(z, z, mode=6, stats=stz))
    timed("pinv 384^3 multi3", lambda: K.gemm(z, z, more=[(z, z), (z.transpose(-1, -2), z.transpose(-1, -2))], out_bf16=zo, res=z, res2=z))
    sys.exit(0)
# s3: [m, n] rows = landmarks, columns = tokens
ql, k = heads(lm, 0), heads(qkv, E)
o16 = torch.empty(B, hd, m, n, device="cuda", dtype=torch.bfloat16)
st = K.softmax_stats((B, hd), m, n, "cuda")
nel = B * hd * n * m
timed("s3 strided normal bf16", lambda: K.gemm(ql, k, out_bf16=o16, alpha=0.1), nel * 2)
timed("s3 strided ROWSTATS", lambda: K.gemm(ql, k, alpha=0.1, mode=K.GEMM_ROWSTATS, stats=st))
timed("s3 strided SOFTMAX bf16", lambda: K.gemm(ql, k, alpha=0.1, mode=K.GEMM_SOFTMAX, stats=st, out_bf16=o16), nel * 2)
timed("s3 strided ROWDOT", lambda: K.gemm(ql, k, mode=K.GEMM_ROWDOT, stats=st, res=o16), nel * 2)
timed("s3 strided SOFTMAX_BWD", lambda: K.gemm(ql, k, alpha=0.1, mode=K.GEMM_SOFTMAX_BWD, stats=st, res=o16, out_bf16=o16), nel * 4)
