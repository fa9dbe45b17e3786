"""GPU bring-up diagnostic for mirror_gemm_bf16 (run under gpurun).

Runs groups of cases in separate subprocesses (a hang or sticky CUDA error in
one group must not hide the others) and prints per-case error summaries that
are useful when debugging blind: max error, where it is, and which 32x32
blocks of the output are wrong.
    python tools/gemm_diag.py            # all groups
    python tools/gemm_diag.py GROUP      # one group in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["basic", "majors", "cluster3", "batched", "epilogue", "splitk", "perf"]


def ref_gemm(a, b, alpha=1.0, bias=None, act=0, res=None, gamma=1.0, beta=0.0, old=None):
    import torch
    v = alpha * torch.matmul(a.float(), b.float().transpose(-1, -2))
    if bias is not None:
        v = v + bias
    if act == 1:
        v = torch.relu(v)
    elif act == 2:
        v = torch.nn.functional.gelu(v)
    if res is not None:
        v = v + gamma * res.float()
    if beta != 0.0:
        v = v + beta * old
    return v


def report(name, got, want, tol=2e-2):
    import torch
    got, want = got.float(), want.float()
    err = (got - want).abs()
    scale = want.abs().max().item() + 1e-6
    mx = err.max().item()
    ok = mx <= tol * scale and bool(torch.isfinite(got).all())
    msg = f"[{'OK ' if ok else 'BAD'}] {name}: max|err|={mx:.3e} (ref max {scale:.3e})"
    if not ok:
        e2 = err.reshape(-1, err.shape[-2], err.shape[-1])
        bi = int(e2.flatten(1).max(1).values.argmax())
        e = e2[bi]
        r, c = divmod(int(e.argmax()), e.shape[1])
        msg += f" worst at batch {bi} row {r} col {c}: got {got.reshape(e2.shape)[bi, r, c]:.4f} want {want.reshape(e2.shape)[bi, r, c]:.4f}"
        R, Cc = e.shape
        rows = []
        for r0 in range(0, min(R, 256), 32):
            rows.append("".join("x" if e[r0:r0 + 32, c0:c0 + 32].max() > tol * scale else "." for c0 in range(0, min(Cc, 512), 32)))
        msg += "\n    bad 32x32 blocks (rows down, cols across):\n    " + "\n    ".join(rows)
        nz = (got.reshape(e2.shape)[bi] == 0).float().mean().item()
        msg += f"\n    fraction of exact zeros in output: {nz:.3f}"
    print(msg, flush=True)
    return ok


def run_group(group):
    import torch
    from mirror_b200 import kernels as K
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    bf = lambda *s: rn(*s).to(torch.bfloat16)
    ok = True

    def simple(name, M, N, K_, a_t=False, b_t=False, **kw):
        a = bf(K_, M).t() if a_t else bf(M, K_)
        b = bf(K_, N).t() if b_t else bf(N, K_)
        o32 = torch.full((M, N), float("nan"), device=dev)
        o16 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        K.gemm(a, b, out_f32=o32, out_bf16=o16, **kw)
        torch.cuda.synchronize()
        want = ref_gemm(a, b, **{k: v for k, v in kw.items() if k in ("alpha", "bias", "act", "res", "gamma")})
        r1 = report(name + " f32", o32, want, 2e-3)
        r2 = report(name + " bf16", o16, want, 1.2e-2)
        return r1 and r2

    if group == "basic":
        ok &= simple("NT 128x128x64", 128, 128, 64)
        ok &= simple("NT 128x256x64", 128, 256, 64)
        ok &= simple("NT 256x256x256", 256, 256, 256)
        ok &= simple("NT 512x768x768", 512, 768, 768)
        ok &= simple("NT 300x200x72 (ragged)", 300, 200, 72)
        ok &= simple("NT 64x3000x768 (small M, odd N)", 64, 3000, 768)
        ok &= simple("NT 2304x384x96", 2304, 384, 96)
        ok &= simple("NT 100x40x24", 100, 40, 24)
        ok &= simple("NT 4096x2304x768 (multi-tile persistent)", 4096, 2304, 768)
    elif group == "majors":
        for (at, bt, nm) in ((False, True, "NN"), (True, False, "TT"), (True, True, "TN")):
            ok &= simple(f"{nm} 128x128x64", 128, 128, 64, at, bt)
            ok &= simple(f"{nm} 256x256x256", 256, 256, 256, at, bt)
            ok &= simple(f"{nm} 384x96x2304", 384, 96, 2304, at, bt)
            ok &= simple(f"{nm} 300x200x72", 304, 200, 72, at, bt)
            ok &= simple(f"{nm} 768x768x4096", 768, 768, 4096, at, bt)
    elif group == "cluster3":
        # 384-row problems run on 3-CTA clusters with TMA-multicast B (the Moore-Penrose / landmark matrices)
        for (at, bt, nm) in ((False, False, "NT"), (False, True, "NN"), (True, False, "TT"), (True, True, "TN")):
            for (Bt, M, N, K_) in ((5, 384, 384, 384), (3, 384, 2304, 96), (2, 304, 192, 200), (70, 384, 384, 384)):
                a = bf(Bt, K_, M).transpose(-1, -2) if at else bf(Bt, M, K_)
                b = bf(Bt, K_, N).transpose(-1, -2) if bt else bf(Bt, N, K_)
                r = bf(Bt, M, N)
                o32 = torch.full((Bt, M, N), float("nan"), device=dev)
                o16 = torch.zeros(Bt, M, N, device=dev, dtype=torch.bfloat16)
                K.gemm(a, b, out_f32=o32, out_bf16=o16, alpha=0.25, res=r, gamma=1.0, diag=1.0 if M == N else 0.0)
                torch.cuda.synchronize()
                want = 0.25 * a.float() @ b.float().transpose(-1, -2) + r.float()
                if M == N:
                    want = want + torch.eye(M, device=dev)
                ok &= report(f"{nm} batch{Bt} {M}x{N}x{K_} f32", o32, want, 2e-3)
                ok &= report(f"{nm} batch{Bt} {M}x{N}x{K_} bf16", o16, want, 1.2e-2)
    elif group == "batched":
        B, n, E, h = 3, 512, 768, 8
        d, m = E // h, 128
        qkv = bf(B, n, 3 * E)
        lm = bf(B, m, 2 * E)
        q = qkv[:, :, :E].reshape(B, n, h, d).permute(0, 2, 1, 3)          # [B,h,n,d] view
        kl = lm[:, :, E:].reshape(B, m, h, d).permute(0, 2, 1, 3)          # [B,h,m,d]
        v = qkv[:, :, 2 * E:].reshape(B, n, h, d).permute(0, 2, 1, 3)
        s1 = torch.empty(B, h, n, m, device=dev)
        K.gemm(q, kl, out_f32=s1, alpha=d ** -0.5)
        torch.cuda.synchronize()
        ok &= report("batched sim1 q@kl^T (strided heads)", s1, d ** -0.5 * q.float() @ kl.float().transpose(-1, -2), 2e-3)
        a3 = torch.softmax(rn(B, h, m, n), -1).to(torch.bfloat16)
        kv = torch.empty(B, h, m, d, device=dev, dtype=torch.bfloat16)
        K.gemm(a3, v.transpose(-1, -2), out_bf16=kv)
        torch.cuda.synchronize()
        ok &= report("batched a3@v (B operand MN-major strided)", kv, a3.float() @ v.float(), 1.2e-2)
        a1 = torch.softmax(rn(B, h, n, m), -1).to(torch.bfloat16)
        w = bf(B, h, m, d)
        out = torch.zeros(B, n, E, device=dev, dtype=torch.bfloat16)
        rc = bf(B, n, E)
        outv = out.reshape(B, n, h, d).permute(0, 2, 1, 3)
        K.gemm(a1, w.transpose(-1, -2), out_bf16=outv, res=rc.reshape(B, n, h, d).permute(0, 2, 1, 3))
        torch.cuda.synchronize()
        want = (a1.float() @ w.float()).permute(0, 2, 1, 3).reshape(B, n, E) + rc.float()
        ok &= report("batched a1@w -> head-merged output + bf16 residual", out, want, 1.2e-2)
        da1 = torch.empty(B, h, m, d, device=dev)
        K.gemm(a1.transpose(-1, -2), outv.transpose(-1, -2), out_f32=da1)   # a1^T @ out : TN batched
        torch.cuda.synchronize()
        ok &= report("batched TN a1^T@dOut", da1, a1.float().transpose(-1, -2) @ outv.float(), 3e-3)
    elif group == "epilogue":
        M, N, K_ = 384, 512, 256
        bias = rn(N)
        ok &= simple("bias+relu", M, N, K_, bias=bias, act=1)
        ok &= simple("bias+gelu alpha", M, N, K_, bias=bias, act=2, alpha=0.37)
        ok &= simple("res f32 gamma", M, N, K_, res=rn(M, N), gamma=-2.5)
        ok &= simple("res bf16 gamma alpha", M, N, K_, res=bf(M, N), gamma=7.0, alpha=-1.0)
        ok &= simple("ragged N bias+relu", 200, 72, 64, bias=rn(72), act=1)
        a, b = bf(M, K_), bf(N, K_)
        o = rn(M, N)
        old = o.clone()
        K.gemm(a, b, out_f32=o, beta=1.0, alpha=0.5)
        torch.cuda.synchronize()
        ok &= report("beta accumulate", o, ref_gemm(a, b, alpha=0.5, beta=1.0, old=old), 2e-3)
        # dropout: tensor-core kernel and SIMT kernel share the hash -> identical masks
        o1, o2 = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
        K.gemm(a, b, out_f32=o1, drop_p=0.1, drop_seed=1234, bias=bias)
        K.gemm(a, b, out_f32=o2, drop_p=0.1, drop_seed=1234, bias=bias, simt=True)
        torch.cuda.synchronize()
        ok &= report("dropout vs simt kernel", o1, o2, 2e-3)
        frac = (o1 == 0).float().mean().item()
        print(f"     dropout zero fraction {frac:.4f} (want ~0.10)")
        ok &= abs(frac - 0.1) < 0.01
        # column-offset output views (non 16B-aligned fall back to scalar stores)
        big = torch.zeros(M, N + 8, device=dev)
        K.gemm(a, b, out_f32=big[:, 3:3 + N])
        torch.cuda.synchronize()
        ok &= report("unaligned output view", big[:, 3:3 + N], ref_gemm(a, b), 2e-3)
        ok &= float(big[:, :3].abs().sum() + big[:, 3 + N:].abs().sum()) == 0.0
    elif group == "splitk":
        for (M, N, K_) in ((768, 768, 8192), (256, 2304, 4000), (128, 128, 64 * 5)):
            a, b = bf(K_, M).t(), bf(K_, N).t()
            o = torch.zeros(M, N, device=dev)
            K.gemm(a, b, out_f32=o, split_k=8)
            torch.cuda.synchronize()
            ok &= report(f"split-K TN {M}x{N}x{K_}", o, ref_gemm(a, b), 2e-3)
    elif group == "perf":
        for (Bt, nm) in ((512, "NN"), (512, "NT")):
            a = bf(Bt, 384, 384)
            b = bf(Bt, 384, 384)
            bb = b.transpose(-1, -2) if nm == "NN" else b
            o = torch.empty(Bt, 384, 384, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                K.gemm(a, bb, out_bf16=o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                K.gemm(a, bb, out_bf16=o)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"[perf] {nm} batched 512 x 384^3 bf16-out: {ms:.3f} ms = {2 * Bt * 384 ** 3 / ms / 1e9:.0f} TFLOP/s", flush=True)
        for (nm, M, N, K_, at, bt) in (("NT qkv-like", 147456 // 4, 2304, 768, False, False),
                                       ("NT square 8192", 8192, 8192, 8192, False, False),
                                       ("NN 8192", 8192, 8192, 8192, False, True),
                                       ("TN wgrad", 768, 2304, 36864, True, True)):
            a = bf(K_, M).t() if at else bf(M, K_)
            b = bf(K_, N).t() if bt else bf(N, K_)
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            o32 = torch.zeros(M, N, device=dev)
            sk = 8 if nm.startswith("TN") else 1
            run = (lambda: K.gemm(a, b, out_f32=o32, split_k=sk)) if sk > 1 else (lambda: K.gemm(a, b, out_bf16=o))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            t0 = time.time()
            c = torch.matmul(a, b.t())
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                c = torch.matmul(a, b.t())
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1) / 10
            print(f"[perf] {nm} {M}x{N}x{K_}: {ms:.3f} ms = {2 * M * N * K_ / ms / 1e9:.0f} TFLOP/s   (torch.matmul {ms_t:.3f} ms = {2 * M * N * K_ / ms_t / 1e9:.0f} TFLOP/s)", flush=True)
    print(f"GROUP {group}: {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run_group(sys.argv[1]) else 1)
    bad = []
    for grp in GROUPS:
        print(f"===== {grp} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), grp], timeout=240)
            if r.returncode != 0:
                bad.append(grp)
        except subprocess.TimeoutExpired:
            print(f"GROUP {grp}: TIMEOUT (hang)", flush=True)
            bad.append(grp)
    print("FAILED GROUPS:", bad)
    sys.exit(1 if bad else 0)
