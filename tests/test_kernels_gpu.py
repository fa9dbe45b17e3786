"""GPU: every C-ABI kernel against its CPU re-statement (tests/emu_backend.py) on the same seeded inputs,
including ragged sizes, padded layouts and strided views.  All calls go through the C ABI (ctypes)."""
import copy

import pytest
import torch

import emu_backend
from mirror_b200 import kernels as K

pytestmark = pytest.mark.gpu
EMU = emu_backend.Emu()
BF16, F32 = torch.bfloat16, torch.float32


def rn(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(s))
    return torch.randn(*s, generator=g) * scale


def to_dev(x):
    if torch.is_tensor(x):
        if x._base is not None or not x.is_contiguous():  # rebuild views on the device copy of the base storage
            base = x._base if x._base is not None else x
            db = base.detach().clone().cuda()
            return torch.as_strided(db, x.shape, x.stride(), x.storage_offset())
        return x.detach().clone().cuda()
    if isinstance(x, (list, tuple)):
        return type(x)(to_dev(v) for v in x)
    return x


def close(a, b, tol, name=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = float(b.abs().max()) + 1e-6
    err = float((a - b).abs().max())
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


def both(op, args, kwargs=None, tol=1e-5, check_args=()):
    """run `op` on the emulation (CPU) and on the device; compare return values and the listed in-place args."""
    kwargs = kwargs or {}
    a_cpu = copy.deepcopy(args)
    a_gpu = to_dev(args)
    r_cpu = getattr(EMU, op)(*a_cpu, **kwargs)
    r_gpu = getattr(K, op)(*a_gpu, **kwargs)
    torch.cuda.synchronize()
    rc = r_cpu if isinstance(r_cpu, (tuple, list)) else (r_cpu,)
    rg = r_gpu if isinstance(r_gpu, (tuple, list)) else (r_gpu,)
    for i, (c, g) in enumerate(zip(rc, rg)):
        if torch.is_tensor(c) and c.numel() > 0 and op not in ("pinv_init",) or (op == "pinv_init" and i < 1):
            close(g, c, tol if c.dtype != BF16 else max(tol, 8e-3), f"{op} ret{i}")
    for i in check_args:
        close(a_gpu[i], a_cpu[i], tol if a_cpu[i].dtype != BF16 else max(tol, 8e-3), f"{op} arg{i}")
    return r_gpu, a_gpu


def test_device_supported():
    from mirror_b200 import _lib
    assert _lib.lib().mirror_device_supported() == 1


@pytest.mark.parametrize("rows,cols,cols_out", [(37, 100, 104), (64, 768, 768), (5, 10234, 10240)])
def test_cast(rows, cols, cols_out):
    both("cast_bf16", (rn(rows, cols),), {"cols_out": cols_out}, tol=1e-6)


def test_cast_strided_rows():
    base = rn(6, 11, 96)
    both("cast_bf16", (base[:, 0, :],), {"cols_out": 96}, tol=1e-6)


@pytest.mark.parametrize("stack,order", [(False, 0), (False, 1), (True, 0), (True, 1)])
def test_cast_split3(stack, order):
    both("cast_split3", (rn(13, 77), 16, 80, stack, order), tol=1e-6)
    both("cast_split3", (rn(13, 72), 16, 80, stack, order), tol=1e-6)  # multiples of 8: the 16-byte vector kernel


def test_split3_product_is_fp32_grade():
    a, b = rn(64, 300).cuda(), rn(48, 300, seed=1).cuda()
    out = torch.empty(64, 48, device="cuda")
    K.gemm(K.cast_split3(a, 64, 304, False, 0), K.cast_split3(b, 48, 304, False, 1), out_f32=out)
    ref = a.double() @ b.double().t()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 2e-5
    plain = torch.empty(64, 48, device="cuda")
    K.gemm(K.cast_bf16(a, 304)[:, :300], K.cast_bf16(b, 304)[:, :300], out_f32=plain)
    assert float((plain.double() - ref).abs().max()) > 20 * float((out.double() - ref).abs().max())


def test_copy_rows_axpy():
    base = rn(5, 9, 64)
    both("copy_rows_", (base[:, 0, :], torch.zeros(5, 64)), check_args=(1,))
    both("axpy_", (rn(1000), rn(1000, seed=3)), {"alpha": 0.3}, check_args=(0,))


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("C", [40, 768])  # 768: the float4 warp-per-row kernel
def test_act_fwd_bwd(act, p, C):
    pre = rn(7, 33, C)
    both("act_fwd", (pre, act, p, 99), {"want_bf16": True, "want_f32": True}, tol=2e-6)
    dy = rn(7, 33, C, seed=5)
    out32 = torch.zeros(7, 40, C)[:, :33, :]  # padded destination view
    out16 = torch.zeros(7, 33, C, dtype=BF16)
    a_gpu = to_dev((dy, pre, out16, out32))
    K.act_bwd(a_gpu[0], a_gpu[1], act, p, 99, out16=a_gpu[2], out32=a_gpu[3])
    c16, c32 = out16.clone(), out32.clone()
    EMU.act_bwd(dy, pre, act, p, 99, out16=c16, out32=c32)
    close(a_gpu[3], c32, 3e-6, "act_bwd f32")
    close(a_gpu[2], c16, 8e-3, "act_bwd bf16")


@pytest.mark.parametrize("N,add", [(97, 3), (100, 0), (2048, 68)])
def test_wsi_assemble_and_embed_bwd(N, add):
    B, E = 3, 64
    S = 1 + N + add
    h = rn(B, S, E)
    both("wsi_assemble_fwd", (h, rn(E, seed=2), N, add), check_args=(0,))
    both("wsi_embed_bwd", (rn(B, S, E, seed=4), h, N, add, torch.zeros(E)), tol=1e-5, check_args=(4,))


@pytest.mark.parametrize("N,keep", [(50, 12), (2048, 512), (768, 192), (1, 0)])
def test_rank_mask(N, keep):
    noise = torch.rand(4, N, generator=torch.Generator().manual_seed(N))
    (m, _) = both("rank_mask", (noise, keep), tol=0)
    assert int(m.sum()) == 4 * (N - keep)


@pytest.mark.parametrize("N,keep", [(1024, 256), (2048, 512), (4099, 1024), (16384, 4096), (3000, 0), (3000, 3000), (2048, 1), (2048, 2047)])
def test_rank_mask_radix_select_long_rows(N, keep):
    """N >= 1024 takes the O(N) radix-select kernel: bit-equal to the stable double argsort, with heavy ties (quantised noise:
    the tie quota and the index order of the equal elements decide), negative values and the keep = 0 / N edges"""
    g = torch.Generator().manual_seed(N + keep)
    cases = [torch.rand(3, N, generator=g), torch.randint(0, 7, (3, N), generator=g).float() / 7.0, torch.randn(3, N, generator=g),
             torch.zeros(2, N)]
    for noise in cases:
        m, _ = both("rank_mask", (noise, keep), tol=0)
        assert int(m.sum()) == noise.shape[0] * (N - keep)


def test_rank_mask_ties_are_stable():
    noise = torch.zeros(2, 64)
    m, _ = both("rank_mask", (noise, 16), tol=0)
    assert torch.equal(m.cpu()[0], (torch.arange(64) >= 16).float())


@pytest.mark.parametrize("first,E,tok_stride", [(1, 48, 1), (0, 1, 0), (1, 192, 1)])  # 192: float4 warp-per-row kernel
def test_mask_pos(first, E, tok_stride):
    B, T = 3, 41
    mask = (torch.rand(B, T - first, generator=torch.Generator().manual_seed(1)) > 0.3).float()
    tok = rn(E if tok_stride else 1, seed=7)
    both("mask_pos_fwd_", (rn(B, T, E), mask, tok, tok_stride, rn(T, E, seed=8), first), check_args=(0,))
    both("mask_pos_bwd", (rn(B, T, E, seed=9), mask, torch.zeros_like(tok), tok_stride, torch.zeros(T, E), first), tol=1e-5,
         check_args=(0, 2, 4))


def test_landmark():
    B, m, seg, E = 2, 24, 3, 48
    qkv = rn(B, m * seg, 3 * E).to(BF16)
    both("landmark_fwd", (qkv, m, seg), tol=8e-3)


def test_gemm_row_broadcast_residual():
    # dq = ds1 @ kl + dql[t // seg] / seg  (landmark-mean backward fused into the GEMM epilogue)
    Bt, n, m, d, seg = 2, 288, 96, 64, 3
    a, b = rn(Bt, n, m, seed=1).to(BF16), rn(Bt, d, m, seed=2).to(BF16)
    r = rn(Bt, m, d, seed=3)
    o_cpu = torch.zeros(Bt, n, d, dtype=BF16)
    EMU.gemm(a, b, out_bf16=o_cpu, res=r, gamma=1.0 / seg, res_row_div=seg)
    o = torch.zeros(Bt, n, d, device="cuda", dtype=BF16)
    K.gemm(a.cuda(), b.cuda(), out_bf16=o, res=r.cuda(), gamma=1.0 / seg, res_row_div=seg)
    close(o, o_cpu, 8e-3, "row-broadcast residual")
    # bf16 residual: the lean epilogue instantiation (bulk stores) takes the same product
    r16 = r.to(BF16)
    EMU.gemm(a, b, out_bf16=o_cpu, res=r16, gamma=1.0 / seg, res_row_div=seg)
    K.gemm(a.cuda(), b.cuda(), out_bf16=o, res=r16.cuda(), gamma=1.0 / seg, res_row_div=seg)
    close(o, o_cpu, 8e-3, "row-broadcast bf16 residual (lean epilogue)")


def test_colsum():
    x = rn(1000, 200)
    both("colsum_", (x[:, :197], torch.zeros(197)), tol=1e-5, check_args=(1,))
    both("colsum_", (x.to(BF16), torch.ones(200)), tol=1e-5, check_args=(1,))


def test_reparam():
    mu, lv, eps = rn(6, 16), rn(6, 16, seed=1), rn(6, 16, seed=2)
    both("reparam_fwd", (mu, lv, eps), tol=2e-6)
    both("reparam_bwd_", (rn(6, 16, seed=3), lv, eps, torch.zeros(6, 16), torch.ones(6, 16)), tol=2e-6, check_args=(3, 4))


@pytest.mark.parametrize("B,S,E,pad", [(2, 65, 192, 31), (1, 300, 768, 0), (3, 17, 1536, 5)])
def test_layernorm(B, S, E, pad):
    x, g, b = rn(B, S, E, scale=2.0) + 0.5, 1 + 0.1 * rn(E, seed=1), 0.1 * rn(E, seed=2)
    both("layernorm_fwd", (x, g, b, 1e-5), {"n_out": S + pad, "pad": pad, "want_bf16": True, "want_f32": True}, tol=3e-6)
    mean, var = x.mean(-1), x.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    dy = rn(B, S + pad, E, seed=3)
    for add in (None, rn(B, S, E, seed=4)):
        both("layernorm_bwd", (dy, x, g, mean, rstd, pad, torch.zeros(B, S, E), add, torch.zeros(E), torch.zeros(E)), tol=2e-5,
             check_args=(6, 8, 9))
    # bf16 upstream gradient (what the token-sized backward GEMM writes)
    both("layernorm_bwd", (dy.to(BF16), x, g, mean, rstd, pad, torch.zeros(B, S, E), None, torch.zeros(E), torch.zeros(E)), tol=2e-5,
         check_args=(6, 8, 9))


@pytest.mark.parametrize("rows,cols", [(100, 384), (7, 2304), (33, 3000), (5, 16512), (64, 96)])
def test_softmax(rows, cols):
    x = rn(rows, cols, scale=3.0)
    (y16, y32), _ = both("softmax_fwd", (x,), {"want_bf16": True, "want_f32": True}, tol=2e-6)
    y = torch.softmax(x, -1).to(BF16)
    both("softmax_bwd", (y, rn(rows, cols, seed=1)), {"scale": 0.3, "want_bf16": True, "want_f32": True}, tol=3e-6)


def test_l2norm_strided_rows():
    base = rn(5, 9, 96)
    x = base[:, 0, :]
    both("l2norm_fwd", (x, 1e-12), tol=2e-6)
    norm = x.norm(dim=-1).clamp_min(1e-12)
    both("l2norm_bwd", (rn(5, 96, seed=1), x, norm), tol=3e-6)


@pytest.mark.parametrize("B,n,E", [(2, 72, 48), (1, 300, 192), (2, 200, 128), (1, 333, 768)])  # head_dim 16 / 96: mma.sync weight grad
def test_res_conv(B, n, E):
    qkv = rn(B, n, 3 * E).to(BF16)
    w = rn(8, 33, scale=0.2)
    both("res_conv_fwd", (qkv, w), tol=8e-3)
    dout = rn(B, n, E, seed=1).to(BF16)
    both("res_conv_bwd", (dout, qkv, w, torch.zeros(8, 33)), tol=2e-4, check_args=(3,))


def test_pinv_init_and_bwd():
    # rows of a real attn2 all sum to 1 (the arg-max row is then decided by round-off); scale the rows apart so that
    # the arg-max row / column are well defined and CPU and GPU must agree on them
    a2 = torch.softmax(rn(2, 8, 24, 24, scale=2.0), -1) * (1 + 0.2 * torch.rand(2, 8, 24, 1, generator=torch.Generator().manual_seed(3)))
    (z16, scratch), _ = both("pinv_init", (a2,), tol=2e-6)
    z_cpu, s_cpu = EMU.pinv_init(a2)
    g = rn(2, 8, 24, 24, seed=1)
    gx_cpu = torch.ones(2, 8, 24, 24)
    EMU.pinv_init_bwd(g, z_cpu, s_cpu, gx_cpu, True)
    gx = torch.ones(2, 8, 24, 24, device="cuda")
    K.pinv_init_bwd(g.cuda(), z16, scratch, gx, True)
    close(gx, gx_cpu, 1e-5, "pinv_init_bwd")


@pytest.mark.parametrize("B,H,E", [(2, 7, 48), (1, 13, 192)])
def test_ppeg(B, H, E):
    x = rn(B, H * H + 1, E)
    w7, w5, w3 = rn(E, 49, scale=0.1), rn(E, 25, scale=0.1), rn(E, 9, scale=0.1)
    b7, b5, b3 = rn(E, seed=1, scale=0.1), rn(E, seed=2, scale=0.1), rn(E, seed=3, scale=0.1)
    (y, wm), _ = both("ppeg_fwd", (x, w7, w5, w3, b7, b5, b3, H), tol=5e-6)
    wm_cpu = EMU.ppeg_fwd(x, w7, w5, w3, b7, b5, b3, H)[1]
    z = lambda *s: torch.zeros(*s)
    both("ppeg_bwd", (rn(B, H * H + 1, E, seed=5), x, wm_cpu, H, z(E, 49), z(E, 25), z(E, 9), z(E), z(E), z(E)), tol=2e-5,
         check_args=(4, 5, 6, 7, 8, 9))


@pytest.mark.parametrize("B,E", [(3, 192), (2, 768)])
def test_rna_attn(B, E):
    qkv = rn(B, 3 * E)
    both("rna_attn_fwd", (qkv,), tol=3e-6)
    both("rna_attn_bwd", (qkv, rn(B, E, seed=1)), tol=1e-5)


def _nystrom_views(B, E, n, m, seed=0):
    """q/k/v interleaved [B,n,3E] and landmarks [B,m,2E] as the Nystrom layer lays them out; returns head views [B,h,rows,d]"""
    h, d = 8, E // 8
    qkv = (rn(B, n, 3 * E, seed=seed) * 0.7).to(BF16)
    lm = (rn(B, m, 2 * E, seed=seed + 1) * 0.7).to(BF16)
    hv = lambda t, c0: t[:, :, c0:c0 + E].unflatten(-1, (h, d)).permute(0, 2, 1, 3)
    return qkv, lm, hv


@pytest.mark.parametrize("B,E,n,m", [(2, 768, 2304, 384), (1, 768, 768, 384), (3, 192, 192, 96), (2, 192, 96, 96), (1, 384, 960, 192)])
def test_flash_softmax_pv(B, E, n, m):
    """csrc/flash_nystrom.cu forward against the dense re-statement: out = softmax(q k_l^T) W + rc (rows = tokens, keys =
    landmarks, strided head views, residual) and kv = softmax(q_l k^T) v (rows = landmarks, keys = tokens)"""
    h, d = 8, E // 8
    qkv, lm, hv = _nystrom_views(B, E, n, m)
    wv = (rn(B, h, m, d, seed=7) * 0.5).to(BF16)
    rc = (rn(B, n, E, seed=8) * 0.3).to(BF16)
    alpha = d ** -0.5
    for args in (
        (hv(qkv, 0), hv(lm, E), wv, alpha, torch.zeros(B, n, E, dtype=BF16).unflatten(-1, (h, d)).permute(0, 2, 1, 3), hv(rc, 0)),  # K-C
        (hv(lm, 0), hv(qkv, E), hv(qkv, 2 * E), alpha, torch.zeros(B, h, m, d, dtype=BF16), None),                              # K-A
    ):
        a_cpu = copy.deepcopy(args)
        a_gpu = to_dev(args)
        l_cpu = EMU.flash_softmax_pv(*a_cpu)
        l_gpu = K.flash_softmax_pv(*a_gpu)
        torch.cuda.synchronize()
        close(l_gpu, l_cpu, 2e-5, "lse2")
        close(a_gpu[4], a_cpu[4], 8e-3, "out")


@pytest.mark.parametrize("keys,d", [(640, 96), (400, 48)])
def test_flash_softmax_pv_moving_reference(keys, d):
    """The single-pass forward takes exponentials relative to the first key block's row maximum and rescales accumulator and row
    sums only when a later block exceeds it by more than 2^20: keys whose magnitude grows 8x per block of 128 force that path in
    every block (and a shrinking sequence never takes it); both must equal the exact two-pass softmax."""
    B, h, R = 2, 3, 300
    for grow in (8.0, 0.125):
        x = (rn(B, h, R, d, seed=21) * 0.5).to(BF16)
        scale = torch.tensor([grow ** (j // 128) for j in range(keys)]).view(1, 1, keys, 1)
        y = (rn(B, h, keys, d, seed=22) * 0.5 * scale).to(BF16)
        v = rn(B, h, keys, d, seed=23).to(BF16)
        alpha = d ** -0.5
        a_cpu = (x, y, v, alpha, torch.zeros(B, h, R, d, dtype=BF16), None)
        a_gpu = to_dev(a_cpu)
        l_cpu = EMU.flash_softmax_pv(*a_cpu)
        l_gpu = K.flash_softmax_pv(*a_gpu)
        torch.cuda.synchronize()
        assert torch.isfinite(a_gpu[4].float()).all()
        close(l_gpu, l_cpu, 2e-5, "lse2")
        close(a_gpu[4], a_cpu[4], 8e-3, "out")


@pytest.mark.parametrize("B,E,n,m", [(2, 768, 2304, 384), (1, 768, 768, 384), (3, 192, 192, 96), (2, 192, 96, 96), (1, 384, 960, 192)])
def test_flash_bwd(B, E, n, m):
    """csrc/flash_nystrom.cu backward, both orientations, for both Nystrom products: recomputed probabilities, dS through shared
    memory, landmark-mean broadcast residual (row_div), value residual, f32 and bf16 outputs, strided head-slot outputs"""
    h, d = 8, E // 8
    seg = n // m
    qkv, lm, hv = _nystrom_views(B, E, n, m)
    alpha = d ** -0.5
    wv = (rn(B, h, m, d, seed=7) * 0.5).to(BF16)
    do = (rn(B, n, E, seed=9) * 0.2).to(BF16)
    dkv = (rn(B, h, m, d, seed=10) * 0.2).to(BF16)
    dlm16 = (rn(B, m, 2 * E, seed=11) * 0.2).to(BF16)
    dvc = (rn(B, n, E, seed=12) * 0.2).to(BF16)
    q, k, v, ql, kl = hv(qkv, 0), hv(qkv, E), hv(qkv, 2 * E), hv(lm, 0), hv(lm, E)
    # forward statistics on the emulator
    o1 = torch.zeros(B, h, n, d, dtype=BF16)
    lse1 = EMU.flash_softmax_pv(q, kl, wv, alpha, o1)
    dot1 = (hv(do, 0).float() * o1.float()).sum(-1)
    o3 = torch.zeros(B, h, m, d, dtype=BF16)
    lse3 = EMU.flash_softmax_pv(ql, k, v, alpha, o3)
    dot3 = (dkv.float() * o3.float()).sum(-1)
    z16 = lambda *s_: torch.zeros(*s_, dtype=BF16)
    z32 = lambda *s_: torch.zeros(*s_, dtype=F32)
    cases = [
        # attn1, rows = tokens: dq (+ landmark broadcast)
        (q, kl, hv(do, 0), wv, alpha, lse1, dot1, False, (hv(z16(B, n, 3 * E), 0), hv(dlm16, 0), seg, 1.0 / seg), None),
        # attn1, keys = landmarks: dkl part (f32 landmark slot), dW
        (kl, q, wv, hv(do, 0), alpha, lse1, dot1, True, (hv(z32(B, m, 2 * E), E), None, 1, 1.0), (z16(B, h, m, d), None, 1, 1.0)),
        # attn3, rows = landmarks: dql part (f32 landmark slot)
        (ql, k, dkv, v, alpha, lse3, dot3, False, (hv(z32(B, m, 2 * E), 0), None, 1, 1.0), None),
        # attn3, keys = tokens: dk (+ landmark broadcast), dv (+ value residual)
        (k, ql, v, dkv, alpha, lse3, dot3, True, (hv(z16(B, n, 3 * E), E), hv(dlm16, E), seg, 1.0 / seg), (hv(z16(B, n, 3 * E), 2 * E), hv(dvc, 0), 1, 1.0)),
    ]
    for i, args in enumerate(cases):
        a_cpu = copy.deepcopy(args)
        a_gpu = to_dev(args)
        EMU.flash_bwd(*a_cpu)
        K.flash_bwd(*a_gpu)
        torch.cuda.synchronize()
        close(a_gpu[8][0], a_cpu[8][0], 1e-2, f"case {i} out1")
        if args[9] is not None:
            close(a_gpu[9][0], a_cpu[9][0], 1e-2, f"case {i} out2")


def _contrastive_problem(Br, Bc, E, precise, seed=0):
    """bf16 operands as ops.contrastive_operands builds them (on the CPU emulator), unit-norm rows, temperature 0.07."""
    x = torch.nn.functional.normalize(rn(Br, E, seed=seed), dim=-1)
    y = torch.nn.functional.normalize(rn(Bc, E, seed=seed + 1), dim=-1)
    Dp = (E + 63) // 64 * 64
    if precise:
        return EMU.cast_split3(x, Br, Dp, False, 0), EMU.cast_split3(y, Bc, Dp, False, 1), Dp
    return EMU.cast_bf16(x, Dp), EMU.cast_bf16(y, Dp), Dp


@pytest.mark.parametrize("Br,Bc,E,diag0", [(37, 37, 64, 0), (5, 5, 40, 0), (300, 300, 192, 0), (256, 256, 512, 0), (64, 256, 768, 128),
                                           (130, 1000, 96, 700), (1024, 1024, 128, 0)])
@pytest.mark.parametrize("precise", [False, True])
def test_fused_contrastive_kernels(Br, Bc, E, diag0, precise):
    """csrc/contrastive.cu: statistics pass and gradient pass (logits only in TMEM) against the dense re-statement,
    ragged row / column tails, several column blocks and splits, the global-negative layout (Bc > Br, diag0 > 0)"""
    x, y, Dp = _contrastive_problem(Br, Bc, E, precise)
    scale = torch.tensor(1 / 0.07)
    (lse, diag), _ = both("contrastive_stats", (x, y, scale), {"diag0": diag0}, tol=2e-5)
    lse_r, _ = EMU.contrastive_stats(x, y, scale, diag0)
    l = scale * (x.float() @ y.float().T)
    lse_c = torch.logsumexp(torch.cat([l, rn(Bc - Br, Bc, seed=5)], 0), 0) if Bc > Br else torch.logsumexp(l, 0)
    a_r, a_c = torch.rand(Br) / Br, torch.rand(Bc) / Bc
    for ac in (a_c, None):
        both("contrastive_grad", (x, y, Dp, E, precise, 2 * Dp, scale, diag0, lse_r, lse_c, a_r, ac, torch.zeros(())),
             tol=2e-3 if not precise else 3e-5, check_args=(12,))


@pytest.mark.parametrize("B", [37, 256, 2048, 8192])
@pytest.mark.parametrize("sym", [True, False])
def test_fused_contrastive_gradients_vs_fp64(B, sym):
    """InfoNCE through the fused kernels at the C5 sizes (D = 512) against the fp64 formula (losses/info_nce.py:144-164) evaluated
    with torch on the device: loss rel <= 1e-3, gradient rel-L2 <= 1e-2 (north-star tolerances; B > 1024 runs plain bf16 operands)"""
    from mirror_b200.losses import InfoNCE
    g = torch.Generator().manual_seed(B)
    q = torch.randn(B, 512, generator=g).cuda().requires_grad_(True)
    k = torch.randn(B, 512, generator=g).cuda().requires_grad_(True)
    loss = InfoNCE(temperature=0.07, symmetric=sym)(q, k)
    loss.backward()
    qd, kd = q.detach().double().requires_grad_(True), k.detach().double().requires_grad_(True)
    ref = O_info_nce(qd, kd, 0.07, sym)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref)), (float(loss), float(ref))
    for a, b, n in ((q.grad, qd.grad, "dq"), (k.grad, kd.grad, "dk")):
        err = float((a.double() - b).norm() / b.norm())
        assert err <= (1e-2 if B > 1024 else 5e-4), (n, err)


def O_info_nce(q, k, t, sym, reduction="mean"):
    from oracle import mirror_oracle as O
    return O.info_nce(q, k, t, sym, reduction)


@pytest.mark.parametrize("reduction", ["none", "sum", "mean"])
@pytest.mark.parametrize("sym", [True, False])
def test_infonce_reductions(reduction, sym):
    """losses/info_nce.py:155-164 passes `reduction` to F.cross_entropy; per-sample upstream gradients reach the kernels"""
    from mirror_b200.losses import InfoNCE
    g = torch.Generator().manual_seed(3)
    q = torch.randn(70, 96, generator=g).cuda().requires_grad_(True)
    k = torch.randn(70, 96, generator=g).cuda().requires_grad_(True)
    wts = torch.rand(70, generator=g).cuda()
    out = InfoNCE(temperature=0.1, symmetric=sym, reduction=reduction)(q, k)
    ((out * wts).sum() if reduction == "none" else out).backward()
    qd, kd = q.detach().double().requires_grad_(True), k.detach().double().requires_grad_(True)
    ref = O_info_nce(qd, kd, 0.1, sym, reduction)
    ((ref * wts.double()).sum() if reduction == "none" else ref).backward()
    close(out, ref, 1e-5, "loss")
    close(q.grad, qd.grad, 2e-4, "dq")
    close(k.grad, kd.grad, 2e-4, "dk")


@pytest.mark.parametrize("E", [32, 192, 768, 1])  # 192 / 768: warp-per-row float4 kernels; 32 / 1: scalar kernels
def test_masked_mse_strided(E):
    B, T = 3, 20
    big = rn(B, T + 1, E)
    a, b = rn(B, T + 1, E, seed=1)[:, 1:, :], big[:, 1:, :]
    mask = (torch.rand(B, T, generator=torch.Generator().manual_seed(2)) > 0.25).float()
    (out, scratch), _ = both("masked_mse_fwd", (a, b, mask), tol=3e-6)
    s_cpu = EMU.masked_mse_fwd(a, b, mask)[1]
    both("masked_mse_bwd", (a, b, mask, s_cpu, torch.tensor(1.3), 0.5, torch.zeros(B, T, E), False, torch.ones(B, T, E), True), tol=3e-6,
         check_args=(6, 8))


def test_gauss_and_sym_kl():
    B, L, P = 4, 16, 3000
    mu, lv = rn(2 * B, L), rn(2 * B, L, seed=1, scale=0.5)
    both("gauss_kl_fwd", (mu, lv, B), tol=3e-6)
    both("gauss_kl_bwd_", (mu, lv, B, torch.tensor(0.9), 0.1, torch.zeros(2 * B, L), torch.zeros(2 * B, L)), tol=3e-6, check_args=(5, 6))
    scores = rn(2 * B, P)
    both("sym_kl_fwd", (scores, B), tol=5e-6)
    both("sym_kl_bwd", (scores, B, torch.tensor(1.1), 0.2), tol=2e-5)
    both("loss_combine", (rn(5), (0.5, 0.1, 0.1, 0.1, 0.2)), tol=1e-6)


def test_gemm_matches_emulation_with_epilogues():
    a, b = rn(200, 72).to(BF16), rn(136, 72, seed=1).to(BF16)
    for kw in (dict(alpha=0.5, bias=rn(136, seed=2), act=1), dict(diag=1.0, alpha=-1.0), dict(res=rn(200, 136, seed=3), gamma=3.25, alpha=-0.25),
               dict(drop_p=0.1, drop_seed=77, bias=rn(136, seed=2)), dict(res=rn(200, 136, seed=3).to(BF16), gamma=7.0, alpha=-1.0)):
        o_cpu, o16_cpu = torch.zeros(200, 136), torch.zeros(200, 136, dtype=BF16)
        EMU.gemm(a, b, out_f32=o_cpu, out_bf16=o16_cpu, **kw)
        kw_g = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
        o, o16 = torch.zeros(200, 136, device="cuda"), torch.zeros(200, 136, device="cuda", dtype=BF16)
        K.gemm(a.cuda(), b.cuda(), out_f32=o, out_bf16=o16, **kw_g)
        close(o, o_cpu, 1e-5, f"gemm {list(kw)}")
        close(o16, o16_cpu, 8e-3, f"gemm bf16 {list(kw)}")


def test_gemm_broadcast_weight_over_batch():
    B, S, E = 3, 70, 64
    x = rn(B, S + 5, E).to(BF16)
    w = rn(E, E, seed=1).to(BF16)
    out_cpu, out = torch.zeros(B, S, E), torch.zeros(B, S, E, device="cuda")
    EMU.gemm(x[:, 5:, :], w.unsqueeze(0).expand(B, E, E), out_f32=out_cpu)
    xd = x.cuda()
    K.gemm(xd[:, 5:, :], w.cuda().unsqueeze(0).expand(B, E, E), out_f32=out)
    close(out, out_cpu, 1e-5, "broadcast gemm")


def test_gemm_multi_term_mixed_layouts_and_second_residual():
    Bt, M = 3, 384
    mk = lambda seed: rn(Bt, M, M, seed=seed, scale=0.2).to(BF16)
    gF, G1, gq, Em, r = mk(1), mk(2), mk(3), mk(4), mk(5)
    o_cpu, o16_cpu = torch.zeros(Bt, M, M), torch.zeros(Bt, M, M, dtype=BF16)
    T = lambda x: x.transpose(-1, -2)
    kw = dict(alpha=-1.0, gamma=-1.0, gamma2=-4.0)
    EMU.gemm(gF, G1, more=[(gq, Em), (T(Em), T(gq))], out_f32=o_cpu, out_bf16=o16_cpu, res=gF, res2=r, **kw)
    d = lambda x: x.cuda()
    gFd, G1d, gqd, Emd, rd = d(gF), d(G1), d(gq), d(Em), d(r)
    o, o16 = torch.zeros(Bt, M, M, device="cuda"), torch.zeros(Bt, M, M, device="cuda", dtype=BF16)
    K.gemm(gFd, G1d, more=[(gqd, Emd), (T(Emd), T(gqd))], out_f32=o, out_bf16=o16, res=gFd, res2=rd, **kw)
    close(o, o_cpu, 2e-5, "multi f32")
    close(o16, o16_cpu, 8e-3, "multi bf16")
    # six terms with different K (the g_a2 reduction over the Moore-Penrose iterations)
    pairs = [(rn(2, 200, 64 + 8 * i, seed=10 + i).to(BF16), rn(2, 136, 64 + 8 * i, seed=20 + i).to(BF16)) for i in range(6)]
    o_cpu = torch.zeros(2, 200, 136)
    EMU.gemm(pairs[0][0], pairs[0][1], more=pairs[1:], out_f32=o_cpu)
    pd = [(a.cuda(), b.cuda()) for a, b in pairs]
    o = torch.zeros(2, 200, 136, device="cuda")
    K.gemm(pd[0][0], pd[0][1], more=pd[1:], out_f32=o)
    close(o, o_cpu, 2e-5, "six-term")


@pytest.mark.parametrize("rows,cols,kd", [(300, 96, 24), (384, 384, 96), (96, 608, 96), (130, 2048, 64), (2048, 384, 96)])
def test_gemm_fused_softmax_fwd_bwd(rows, cols, kd):
    # softmax(alpha q k^T) and its backward as GEMM epilogue modes, checked against an fp32 softmax of the same product
    Bt, hd, alpha = 2, 3, kd ** -0.5
    a, b = (3 * rn(Bt, hd, rows, kd, seed=1)).to(BF16), rn(Bt, hd, cols, kd, seed=2).to(BF16)
    logits = alpha * a.float() @ b.float().transpose(-1, -2)
    p_ref = torch.softmax(logits, -1)
    ad, bd = a.cuda(), b.cuda()
    st = K.softmax_stats((Bt, hd), rows, cols, "cuda")
    assert st.shape[-2] == K.gemm_nparts(cols)
    K.gemm(ad, bd, alpha=alpha, mode=K.GEMM_ROWSTATS, stats=st)
    p16 = torch.zeros(Bt, hd, rows, cols, device="cuda", dtype=BF16)
    p32 = torch.zeros(Bt, hd, rows, cols, device="cuda")
    K.gemm(ad, bd, alpha=alpha, mode=K.GEMM_SOFTMAX, stats=st, out_bf16=p16, out_f32=p32)
    assert (p32.cpu() - p_ref).abs().max() <= 2e-6 + 1e-5 * p_ref.abs().max()
    close(p16, p_ref, 6e-3, "softmax bf16")
    assert (p32.sum(-1) - 1).abs().max() < 1e-5
    # backward: G = ga gb^T, ds = alpha * P * (G - rowsum(G P)) with P the stored bf16 probabilities
    ga, gb = rn(Bt, hd, rows, 40, seed=3).to(BF16), rn(Bt, hd, cols, 40, seed=4).to(BF16)
    G = ga.float() @ gb.float().transpose(-1, -2)
    P = p16.cpu().float()
    ds_ref = alpha * P * (G - (G * P).sum(-1, keepdim=True))
    st2 = K.softmax_stats((Bt, hd), rows, cols, "cuda")
    K.gemm(ga.cuda(), gb.cuda(), mode=K.GEMM_ROWDOT, stats=st2, res=p16)
    ds = torch.zeros(Bt, hd, rows, cols, device="cuda", dtype=BF16)
    K.gemm(ga.cuda(), gb.cuda(), alpha=alpha, mode=K.GEMM_SOFTMAX_BWD, stats=st2, res=p16, out_bf16=ds)
    close(ds, ds_ref, 6e-3, "softmax bwd")
    # one-pass form: the row dots are handed over (here computed on the host), MIRROR_GEMM_SOFTMAX_BWD_DOT
    ds1 = torch.zeros_like(ds)
    K.gemm(ga.cuda(), gb.cuda(), alpha=alpha, mode=K.GEMM_SOFTMAX_BWD_DOT, stats=(G * P).sum(-1).cuda().contiguous(), res=p16, out_bf16=ds1)
    close(ds1, ds_ref, 6e-3, "softmax bwd (dots given)")
    # and the emulation used by the CPU suite follows the same contract
    st_c = torch.zeros(Bt, hd, rows, K.gemm_nparts(cols), 2)
    EMU.gemm(a, b, alpha=alpha, mode=1, stats=st_c)
    p_c = torch.zeros(Bt, hd, rows, cols)
    EMU.gemm(a, b, alpha=alpha, mode=2, stats=st_c, out_f32=p_c)
    assert (p_c - p_ref).abs().max() < 1e-5


@pytest.mark.parametrize("M,N,Kd,outs", [(300, 96, 64, "both"), (130, 384, 96, "b16"), (384, 192, 384, "f32"), (77, 160, 40, "both")])
def test_gemm_bulk_store_epilogue(M, N, Kd, outs):
    # short-K products hand finished 32x32 chunks to TMA (staging box -> bulk tensor store); row tails are clipped by the
    # tensor map, outputs may be head-strided views.  Checked against the emulation, residual + diag + alpha included.
    Bt, hd = 2, 3
    a, b = rn(Bt, hd, M, Kd, seed=1).to(BF16), rn(Bt, hd, N, Kd, seed=2).to(BF16)
    r = rn(Bt, hd, M, N, seed=3).to(BF16)
    kw = dict(alpha=0.37, diag=1.0, res=r, gamma=-0.5)
    o32c = torch.zeros(Bt, hd, M, N) if outs != "b16" else None
    o16c = torch.zeros(Bt, M, hd * N, dtype=BF16).unflatten(-1, (hd, N)).permute(0, 2, 1, 3) if outs != "f32" else None
    EMU.gemm(a, b, out_f32=o32c, out_bf16=o16c, **kw)
    o32 = torch.full((Bt, hd, M + 1, N), 7.0, device="cuda") if outs != "b16" else None           # one guard row per matrix
    o16s = torch.full((Bt, M + 1, hd * N), 7.0, device="cuda", dtype=BF16) if outs != "f32" else None
    o16 = o16s[:, :M].unflatten(-1, (hd, N)).permute(0, 2, 1, 3) if o16s is not None else None
    kwd = dict(kw, res=r.cuda())
    K.gemm(a.cuda(), b.cuda(), out_f32=o32[:, :, :M] if o32 is not None else None, out_bf16=o16, **kwd)
    if o32 is not None:
        close(o32[:, :, :M], o32c, 2e-5, "bulk-store f32")
        assert (o32[:, :, M] == 7.0).all(), "rows past M must be clipped"
    if o16 is not None:
        close(o16, o16c, 8e-3, "bulk-store bf16")
        assert (o16s[:, M] == 7.0).all(), "rows past M must be clipped"


@pytest.mark.parametrize("B,X,S,E", [(3, 40, 33, 768), (2, 19, 17, 192)])
def test_layernorm_leading_rows_only(B, X, S, E):
    # final norm of the WSI encoder: only the first S of X rows per slide are normalised / receive a gradient
    x, g, b = rn(B, X, E, scale=2.0) + 0.5, 1 + 0.1 * rn(E, seed=1), 0.1 * rn(E, seed=2)
    both("layernorm_fwd", (x, g, b, 1e-5), {"want_bf16": True, "want_f32": True, "rows": S}, tol=3e-6)
    xs = x[:, :S]
    mean, rstd = xs.mean(-1), torch.rsqrt(xs.var(-1, unbiased=False) + 1e-5)
    dy = rn(B, S, E, seed=3)
    for add in (None, rn(B, X, E, seed=4)):
        both("layernorm_bwd", (dy, x, g, mean, rstd, 0, torch.full((B, X, E), 9.0), add, torch.zeros(E), torch.zeros(E)), tol=2e-5,
             check_args=(6, 8, 9))


def test_token_fanout_bwd():
    B, T, E = 3, 21, 192
    full, cls = rn(B, T, E, seed=1), rn(B, E, seed=2)
    tok_store = rn(B, T + 3, E, seed=3)
    for f, c, t in ((full, cls, tok_store[:, 4:, :]), (None, None, tok_store[:, 4:, :]), (full, None, None), (None, cls, None)):
        want = EMU.token_fanout_bwd(f, c, t, B, T, E, "cpu")
        dev = lambda v: None if v is None else v.cuda()
        td = tok_store.cuda()[:, 4:, :] if t is not None else None  # strided view on the device too
        got = K.token_fanout_bwd(dev(f), dev(c), td, B, T, E, "cuda")
        close(got, want, 1e-6, "token_fanout_bwd")


def test_rowdot():
    a, b = rn(3, 5, 77, 96, seed=1).to(BF16), rn(3, 5, 77, 96, seed=2).to(BF16)
    both("rowdot", (a, b), tol=2e-6)
    both("rowdot", (a, b, rn(3, 5, 77, 96, seed=3).to(BF16)), tol=2e-6)


def test_pinv_init_softmax_bwd_fused():
    # pinv_init_bwd + row-softmax backward of attn2 in one pass (m % 32 == 0), against the two-step emulation
    m = 96
    a2 = torch.softmax(rn(2, 3, m, m, scale=2.0), -1) * (1 + 0.2 * torch.rand(2, 3, m, 1, generator=torch.Generator().manual_seed(3)))
    z16, scratch = K.pinv_init(a2.cuda())
    z_cpu, s_cpu = EMU.pinv_init(a2)
    ga2, gz0 = rn(2, 3, m, m, seed=1), rn(2, 3, m, m, seed=2)
    p16 = torch.softmax(rn(2, 3, m, m, seed=4), -1).to(BF16)
    want = EMU.pinv_init_softmax_bwd(ga2, gz0, z_cpu, p16, s_cpu, 0.3)
    got = K.pinv_init_softmax_bwd(ga2.cuda(), gz0.cuda(), z16, p16.cuda(), scratch, 0.3)
    close(got, want, 8e-3, "pinv_init_softmax_bwd")
