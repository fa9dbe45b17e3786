"""Tensor-level front end of the C-ABI kernels (include/mirror_b200.h).

Every function takes torch tensors (device memory owned by PyTorch's caching
allocator), extracts raw pointers / strides and enqueues the kernel on the
current CUDA stream.  There is no CPU or eager fallback: a missing library or a
non-CUDA tensor raises.

(The CPU unit tests replace these functions from the outside --
tests/emu_backend.py monkeypatches a torch re-statement of each entry point
onto this module -- so that the host logic and the hand-written backward passes
can be checked against the oracle without a GPU.  Nothing in this package
knows about that.)
"""
import ctypes

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
LAUNCHES = [0]  # number of kernels enqueued (bench.py reports it as gpu_launches)


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream():
    # torch.cuda.current_stream() costs ~15 us of Python per call (2400 calls per step); the raw query is a C call
    return ctypes.c_void_p(_raw_stream(_cur_device()))


def _cuda(t, dtype=None):
    if not t.is_cuda:
        raise RuntimeError("mirror_b200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t


def _as4(t):
    """View an operand as [b2, b1, rows, cols]."""
    while t.dim() < 4:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise ValueError("gemm operands have at most 2 batch dims")
    return t


def _major(t):
    """(mn_major, ld) of a [.., rows(MN), cols(K)] view."""
    if t.stride(-1) == 1:
        return 0, t.stride(-2)
    if t.stride(-2) == 1:
        return 1, t.stride(-1)
    raise ValueError(f"gemm operand needs a unit stride in one of its last two dims, got {t.stride()}")


def _operands(g, a, b):
    a4, b4 = _as4(_cuda(a, torch.bfloat16)), _as4(_cuda(b, torch.bfloat16))
    B2, B1, M, K = a4.shape
    N = b4.shape[2]
    if b4.shape != (B2, B1, N, K):
        raise ValueError(f"gemm shape mismatch {tuple(a.shape)} x {tuple(b.shape)}")
    g.a, g.b = a4.data_ptr(), b4.data_ptr()
    g.a_mn_major, g.lda = _major(a4)
    g.b_mn_major, g.ldb = _major(b4)
    g.a_bs1, g.a_bs2, g.b_bs1, g.b_bs2 = a4.stride(1), a4.stride(0), b4.stride(1), b4.stride(0)
    g.M, g.N, g.K, g.batch1, g.batch2 = M, N, K, B1, B2
    return B2, B1, M, N


GEMM_NORMAL, GEMM_ROWSTATS, GEMM_SOFTMAX, GEMM_ROWDOT, GEMM_SOFTMAX_BWD, GEMM_SOFTMAX_BWD_DOT = range(6)


def gemm_nparts(n):
    """Partial-statistics slots per row of an N-column product (two per N tile of the tcgen05 kernel)."""
    return int(_lib.fn("mirror_gemm_nparts")(n))


def softmax_stats(batch_shape, m, n, device):
    """Scratch for the two-pass fused softmax GEMMs: [*batch, M, nparts, 2] f32."""
    return torch.empty(*batch_shape, m, gemm_nparts(n), 2, dtype=torch.float32, device=device)


def gemm(a, b, *, out_f32=None, out_bf16=None, alpha=1.0, bias=None, act=ACT_NONE, drop_p=0.0, drop_seed=0,
         res=None, gamma=1.0, beta=0.0, split_k=1, diag=0.0, more=None, res2=None, gamma2=1.0, res_row_div=1, mode=GEMM_NORMAL,
         stats=None, simt=False):
    """C[..,M,N] = epilogue(A[..,M,K] @ B[..,N,K]^T [+ sum of A_t @ B_t^T for (A_t, B_t) in `more`]); see mirror_gemm_args.

    ``a`` / ``b`` are bf16 views whose last two dims are (M,K) / (N,K); either of the two may carry the unit stride, so
    ``x.transpose(-1,-2)`` views give the NN / TN forms without copies.  Up to two leading batch dims.  ``more``: up to
    five further operand pairs with the same M, N and batch dims, accumulated into the same tile (one epilogue pass).
    ``mode`` / ``stats``: fused row-softmax epilogues (GEMM_ROWSTATS .. GEMM_SOFTMAX_BWD), stats = softmax_stats(...).
    """
    nterms = 1 + (len(more) if more else 0)
    terms = (_lib.GemmArgs * nterms)()
    g = terms[0]
    B2, B1, M, N = _operands(g, a, b)
    for t, (at, bt) in enumerate(more or (), start=1):
        if _operands(terms[t], at, bt) != (B2, B1, M, N):
            raise ValueError("gemm: every term must share M, N and the batch dims")
    g.alpha = alpha
    g.bias = _cuda(bias, torch.float32).data_ptr() if bias is not None else None
    g.act, g.drop_p, g.drop_seed = act, drop_p, drop_seed
    if res is not None:
        r4 = _as4(_cuda(res))
        if r4.shape != (B2, B1, (M + res_row_div - 1) // res_row_div, N) or r4.stride(-1) != 1:
            raise ValueError("gemm residual must be [..,ceil(M/res_row_div),N] with unit column stride")
        g.res_row_div = res_row_div
        g.res, g.res_is_bf16 = r4.data_ptr(), int(r4.dtype == torch.bfloat16)
        if r4.dtype not in (torch.bfloat16, torch.float32):
            raise TypeError("residual must be bf16 or f32")
        g.ldr, g.r_bs1, g.r_bs2 = r4.stride(2), r4.stride(1), r4.stride(0)
        if res2 is not None:
            q4 = _as4(_cuda(res2, torch.bfloat16))
            if q4.shape != r4.shape or q4.stride() != r4.stride():
                raise ValueError("gemm res2 must be bf16 with the shape and element strides of res")
            g.res2, g.gamma2 = q4.data_ptr(), gamma2
    elif res2 is not None:
        raise ValueError("gemm res2 needs res")
    g.gamma, g.beta, g.split_k, g.diag = gamma, beta, split_k, diag
    if mode != GEMM_NORMAL:
        _cuda(stats, torch.float32)
        want = (*a.shape[:-2], M) if mode == GEMM_SOFTMAX_BWD_DOT else (*a.shape[:-2], M, gemm_nparts(N), 2)
        if tuple(stats.shape) != want or not stats.is_contiguous():
            raise ValueError("gemm: stats must come from softmax_stats() / rowdot() for this product")
        g.mode, g.stats = mode, stats.data_ptr()
    for o, dt, name in ((out_f32, torch.float32, "32"), (out_bf16, torch.bfloat16, "16")):
        if o is None:
            continue
        o4 = _as4(_cuda(o, dt))
        if o4.shape != (B2, B1, M, N) or o4.stride(-1) != 1:
            raise ValueError(f"gemm output must be [..,M,N]={B2, B1, M, N} with unit column stride, got {tuple(o.shape)}")
        if name == "32":
            g.out_f32, g.ldc32, g.c32_bs1, g.c32_bs2 = o4.data_ptr(), o4.stride(2), o4.stride(1), o4.stride(0)
        else:
            g.out_bf16, g.ldc16, g.c16_bs1, g.c16_bs2 = o4.data_ptr(), o4.stride(2), o4.stride(1), o4.stride(0)
    if nterms > 1:
        if simt:
            raise ValueError("the SIMT cross-check kernel has no multi-term form")
        _lib.check(_lib.fn("mirror_gemm_bf16_multi")(ctypes.cast(terms, ctypes.c_void_p), nterms, _stream()), "gemm_multi")
    else:
        fn = _lib.lib().mirror_gemm_bf16_simt if simt else _lib.lib().mirror_gemm_bf16
        _lib.check(fn(ctypes.byref(g), _stream()), "gemm")
    LAUNCHES[0] += 1


# =====================================================================================
# non-GEMM entry points.  Every function below has a same-named re-statement in
# tests/emu_backend.py (CPU tests only).
# =====================================================================================
import functools


def _op(f):
    """marks a tensor-level entry point (one C-ABI call each); no behaviour of its own"""
    return f


def _p(t, dtype=None):
    if t is None:
        return None
    _cuda(t, dtype)
    return t.data_ptr()


def _call(name, *args, launches=1):
    _lib.check(_lib.fn(name)(*args, _stream()), name)
    LAUNCHES[0] += launches


F32, BF16 = torch.float32, torch.bfloat16


def _contig(t):
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return t


@_op
def cast_bf16(src, cols_out=None):
    """[.., cols] f32 (contiguous, or a 2-D view with unit column stride) -> bf16 [.., cols_out] zero padded."""
    cols = src.shape[-1]
    cols_out = cols_out or cols
    if src.dim() == 2 and src.stride(1) == 1:
        lds = src.stride(0)
    else:
        _contig(src)
        lds = cols
    rows = src.numel() // cols
    dst = torch.empty(*src.shape[:-1], cols_out, device=src.device, dtype=BF16)
    _call("mirror_cast_f32_bf16", _p(src, F32), rows, cols, lds, _p(dst), cols_out, cols_out)
    return dst


@_op
def cast_split3(src, rows_out, cols_out, stack_rows, order):
    """2-D f32 view (unit column stride) -> bf16 split-3 operand: [rows_out, 3*cols_out] (stack_rows=0) or
    [3*rows_out, cols_out] (stack_rows=1); blocks (hi,lo,hi) for order 0, (hi,hi,lo) for order 1."""
    rows, cols = src.shape
    assert src.stride(1) == 1 or cols == 1
    shape = (3 * rows_out, cols_out) if stack_rows else (rows_out, 3 * cols_out)
    dst = torch.empty(shape, device=src.device, dtype=BF16)
    _call("mirror_cast_split3", _p(src, F32), rows, cols, src.stride(0), _p(dst), rows_out, cols_out, int(stack_rows), order)
    return dst


@_op
def copy_rows_(src, dst):
    """dst[r,:] = src[r,:] for 2-D f32 views with unit column stride."""
    rows, cols = src.shape
    assert dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1
    _call("mirror_copy_rows_f32", _p(src, F32), src.stride(0), rows, cols, _p(dst, F32), dst.stride(0))
    return dst


@_op
def axpy_(dst, src, alpha=1.0):
    _contig(dst), _contig(src)
    assert dst.numel() == src.numel()
    _call("mirror_axpy_f32", _p(dst, F32), _p(src, F32), dst.numel(), alpha)
    return dst


@_op
def rowdot(a16, b16, sub16=None):
    """[..., R, C] bf16 (contiguous) -> [..., R] f32 row dots <a, b - sub>."""
    _contig(a16), _contig(b16)
    assert a16.shape == b16.shape and (sub16 is None or (sub16.shape == a16.shape and sub16.is_contiguous()))
    out = torch.empty(a16.shape[:-1], device=a16.device, dtype=F32)
    _call("mirror_rowdot_bf16", _p(a16, BF16), _p(b16, BF16), _p(sub16, BF16), a16.numel() // a16.shape[-1], a16.shape[-1], _p(out))
    return out


@_op
def token_fanout_bwd(d_full, d_cls, d_tok, B, T, E, device):
    """-> [B,T,E] f32 = d_full + (row 0: d_cls) + (rows 1..: d_tok); absent terms are zero (one pass)."""
    out = torch.empty(B, T, E, device=device, dtype=F32)
    bs = ld = 0
    if d_tok is not None:
        assert d_tok.shape == (B, T - 1, E) and d_tok.stride(2) == 1
        bs, ld = d_tok.stride(0), d_tok.stride(1)
    _call("mirror_token_fanout_bwd", _p(_contig(d_full) if d_full is not None else None, F32),
          _p(_contig(d_cls) if d_cls is not None else None, F32), _p(d_tok, F32), bs, ld, B, T, E, _p(out))
    return out


@_op
def act_fwd(pre, act, drop_p=0.0, seed=0, want_bf16=True, want_f32=False):
    _contig(pre)
    o16 = torch.empty_like(pre, dtype=BF16) if want_bf16 else None
    o32 = torch.empty_like(pre) if want_f32 else None
    _call("mirror_act_fwd", _p(pre, F32), pre.numel(), act, drop_p, seed, _p(o16), _p(o32))
    return o16, o32


def _v3(t):
    """(ptr, batch stride, row stride) of a [B,T,C] view with unit stride on C."""
    if t is None:
        return None, 0, 0
    assert t.dim() == 3 and t.stride(2) == 1
    return t.data_ptr(), t.stride(0), t.stride(1)


@_op
def act_bwd(dy, pre, act, drop_p=0.0, seed=0, out16=None, out32=None):
    """dy, pre, outputs: [B,T,C] views with unit stride on C (batch / row strides free)."""
    B, T, C = dy.shape
    for t in (pre, out16, out32):
        assert t is None or tuple(t.shape) == (B, T, C)
    _cuda(dy, F32)
    a, b, c, d = _v3(dy), _v3(pre), _v3(out16), _v3(out32)
    _call("mirror_act_bwd", a[0], a[1], a[2], b[0], b[1], b[2], B, T, C, act, drop_p, seed, c[0], c[1], c[2], d[0], d[1], d[2])


@_op
def wsi_assemble_fwd(h, cls, N, add):
    B, S, E = h.shape
    assert S == 1 + N + add
    _call("mirror_wsi_assemble_fwd", _p(_contig(h), F32), _p(_contig(cls), F32), B, N, add, E)


@_op
def wsi_embed_bwd(dh, h, N, add, dcls):
    """-> dpre16 [B,N,E] (ReLU-masked gradient of the fc1 output incl. the wrap-around rows); dcls accumulated."""
    B, S, E = dh.shape
    dpre = torch.empty(B, N, E, device=dh.device, dtype=BF16)
    _call("mirror_wsi_embed_bwd", _p(_contig(dh), F32), _p(_contig(h), F32), B, N, add, E, _p(dpre), _p(_contig(dcls), F32))
    return dpre


@_op
def rank_mask(noise, keep):
    B, N = noise.shape
    mask = torch.empty(B, N, device=noise.device, dtype=F32)
    _call("mirror_rank_mask", _p(_contig(noise), F32), B, N, keep, _p(mask))
    return mask


@_op
def mask_pos_fwd_(r, mask, tok, tok_stride, pos, first):
    """r: [B,T,E] f32, modified in place."""
    B, T, E = r.shape
    _call("mirror_mask_pos_fwd", _p(_contig(r), F32), _p(_contig(mask), F32), _p(tok, F32), tok_stride, _p(_contig(pos), F32), B, T, E, first)
    return r


@_op
def mask_pos_bwd(dy, mask, dtok, tok_stride, dpos, first):
    """-> dr [B,T,E] = dy with the masked slots zeroed (new tensor); dtok / dpos are ACCUMULATED."""
    B, T, E = dy.shape
    dr = torch.empty_like(dy)
    _call("mirror_mask_pos_bwd", _p(_contig(dy), F32), _p(_contig(mask), F32), _p(dr), _p(dtok, F32), tok_stride, _p(_contig(dpos), F32),
          B, T, E, first)
    return dr


@_op
def landmark_fwd(qkv, m, seg):
    B, n, E3 = qkv.shape
    lm = torch.empty(B, m, 2 * (E3 // 3), device=qkv.device, dtype=BF16)
    _call("mirror_landmark_fwd", _p(_contig(qkv), BF16), _p(lm), B, n, m, seg, E3 // 3)
    return lm


@_op
def colsum_(x, out):
    """out[c] += sum_r x[r,c]; x is a 2-D view with unit column stride (bf16 or f32)."""
    rows, cols = x.shape
    assert x.stride(1) == 1 and out.numel() == cols
    _call("mirror_colsum", _p(x), int(x.dtype == BF16), rows, cols, x.stride(0), _p(out, F32))
    return out


@_op
def reparam_fwd(mu, logvar, eps):
    z = torch.empty_like(mu)
    _call("mirror_reparam_fwd", _p(_contig(mu), F32), _p(_contig(logvar), F32), _p(_contig(eps), F32), mu.numel(), None, _p(z))
    return z


@_op
def reparam_bwd_(dz, logvar, eps, dmu, dlogvar):
    _call("mirror_reparam_bwd", _p(_contig(dz), F32), _p(logvar, F32), _p(eps, F32), dz.numel(), _p(dmu, F32), _p(dlogvar, F32))


@_op
def layernorm_fwd(x, gamma, beta, eps, n_out=None, pad=0, want_bf16=True, want_f32=False, rows=None):
    """x: [B,X,E] f32, first S = ``rows`` (default X) rows per slide -> (y16 [B,n_out,E], y32 [B,n_out,E], mean [B,S], rstd [B,S])."""
    B, X, E = x.shape
    S = rows or X
    n_out = n_out or S
    y16 = torch.empty(B, n_out, E, device=x.device, dtype=BF16) if want_bf16 else None
    y32 = torch.empty(B, n_out, E, device=x.device, dtype=F32) if want_f32 else None
    mean = torch.empty(B, S, device=x.device, dtype=F32)
    rstd = torch.empty(B, S, device=x.device, dtype=F32)
    _call("mirror_layernorm_fwd", _p(_contig(x), F32), _p(gamma, F32), _p(beta, F32), eps, B, S, X, E, n_out, pad, _p(y16), _p(y32),
          _p(mean), _p(rstd))
    return y16, y32, mean, rstd


@_op
def layernorm_bwd(dy, x, gamma, mean, rstd, pad, dx, add, dgamma, dbeta):
    """dy: [B,n_out,E] f32 or bf16; x, dx: [B,X,E] f32, S = mean.shape[1] <= X rows per slide took part in the forward;
    dx = (add or 0) + LN gradient (rows >= S: just add / 0); dgamma/dbeta accumulate."""
    B, X, E = x.shape
    if dy.dtype not in (F32, BF16):
        raise TypeError("layernorm_bwd: dy must be f32 or bf16")
    _call("mirror_layernorm_bwd", _p(_contig(dy)), int(dy.dtype == BF16), _p(_contig(x), F32), _p(gamma, F32), _p(mean), _p(rstd), B, mean.shape[1], X, E,
          dy.shape[1], pad, _p(_contig(dx), F32), _p(add, F32), _p(dgamma, F32), _p(dbeta, F32))


@_op
def softmax_fwd(x, want_bf16=True, want_f32=False):
    _contig(x)
    cols = x.shape[-1]
    y16 = torch.empty_like(x, dtype=BF16) if want_bf16 else None
    y32 = torch.empty_like(x) if want_f32 else None
    _call("mirror_softmax_fwd", _p(x, F32), x.numel() // cols, cols, _p(y16), _p(y32))
    return y16, y32


@_op
def softmax_bwd(y16, dy, scale=1.0, want_bf16=True, want_f32=False):
    _contig(y16), _contig(dy)
    cols = dy.shape[-1]
    d16 = torch.empty_like(dy, dtype=BF16) if want_bf16 else None
    d32 = torch.empty_like(dy) if want_f32 else None
    _call("mirror_softmax_bwd", _p(y16, BF16), _p(dy, F32), dy.numel() // cols, cols, scale, _p(d16), _p(d32))
    return d16, d32


@_op
def l2norm_fwd(x, eps):
    """x: [rows, cols] f32 view (unit column stride) -> (y32 contiguous, norm[rows])."""
    rows, cols = x.shape
    assert x.stride(1) == 1
    y = torch.empty(rows, cols, device=x.device, dtype=F32)
    norm = torch.empty(rows, device=x.device, dtype=F32)
    _call("mirror_l2norm_fwd", _p(x, F32), x.stride(0), rows, cols, eps, None, _p(y), cols, _p(norm))
    return y, norm


@_op
def l2norm_bwd(dy, x, norm):
    rows, cols = x.shape
    dx = torch.empty(rows, cols, device=x.device, dtype=F32)
    _call("mirror_l2norm_bwd", _p(_contig(dy), F32), cols, _p(x, F32), x.stride(0), _p(norm), rows, cols, _p(dx), cols, 0)
    return dx


@_op
def res_conv_fwd(qkv, w):
    B, n, E3 = qkv.shape
    out = torch.empty(B, n, E3 // 3, device=qkv.device, dtype=BF16)
    _call("mirror_res_conv_fwd", _p(_contig(qkv), BF16), _p(_contig(w), F32), B, n, E3 // 3, _p(out))
    return out


@_op
def res_conv_bwd(dout16, qkv, w, dw):
    """-> dv16 [B,n,E] = conv^T(dout) (gradient w.r.t. the value slot); dw [8,33] is accumulated."""
    B, n, E3 = qkv.shape
    dv = torch.empty(B, n, E3 // 3, device=qkv.device, dtype=BF16)
    _call("mirror_res_conv_bwd", _p(_contig(dout16), BF16), _p(_contig(qkv), BF16), _p(_contig(w), F32), B, n, E3 // 3,
          _p(dv), _p(_contig(dw), F32), launches=2)
    return dv


@_op
def pinv_init(a2, z16=None, scratch=None):
    """a2: [..,m,m] f32 -> (z16, scratch) with z0 = a2^T / (max rowsum * max colsum) in bf16, the maxima taken over ALL matrices of
    a2 (the reference's batch-global scale).  z16 / scratch: optional preallocated outputs (slices of per-slide buffers)."""
    m = a2.shape[-1]
    BH = a2.numel() // (m * m)
    z16 = torch.empty_like(a2, dtype=BF16) if z16 is None else z16
    scratch = torch.empty(8, device=a2.device, dtype=F32) if scratch is None else scratch
    _call("mirror_pinv_init", _p(_contig(a2), F32), BH, m, _p(_contig(scratch)), None, _p(_contig(z16)), launches=2)
    return z16, scratch


@_op
def pinv_init_bwd(gz0, z0_16, scratch, gx, accumulate):
    m = gz0.shape[-1]
    BH = gz0.numel() // (m * m)
    _call("mirror_pinv_init_bwd", _p(_contig(gz0), F32), _p(_contig(z0_16), BF16), BH, m, _p(scratch), _p(_contig(gx), F32),
          int(accumulate), launches=2)


@_op
def pinv_init_softmax_bwd(ga2, gz0, z0_16, a2_16, scratch, scale):
    """-> ds2 bf16 = softmax_bwd(a2, ga2 + pinv_init_bwd(gz0)) * scale in one pass (ga2 is not modified)."""
    m = gz0.shape[-1]
    BH = gz0.numel() // (m * m)
    ds = torch.empty_like(a2_16)
    _call("mirror_pinv_init_softmax_bwd", _p(_contig(ga2), F32), _p(_contig(gz0), F32), _p(_contig(z0_16), BF16), _p(_contig(a2_16), BF16),
          BH, m, _p(scratch), scale, _p(ds), launches=2)
    return ds


@_op
def ppeg_fwd(x, w7, w5, w3, b7, b5, b3, H):
    B, S, E = x.shape
    assert S == H * H + 1
    wm = torch.empty(49, E, device=x.device, dtype=F32)
    bm = torch.empty(E, device=x.device, dtype=F32)
    y = torch.empty_like(x)
    _call("mirror_ppeg_fwd", _p(_contig(x), F32), _p(_contig(w7), F32), _p(_contig(w5), F32), _p(_contig(w3), F32), _p(b7, F32),
          _p(b5, F32), _p(b3, F32), B, H, E, _p(wm), _p(bm), _p(y), launches=2)
    return y, wm


@_op
def ppeg_bwd(dy, x, wm, H, dw7, dw5, dw3, db7, db5, db3):
    """returns dx; weight/bias gradients are ACCUMULATED into dw*, db*."""
    B, S, E = x.shape
    dx = torch.empty_like(x)
    dwm = torch.empty(49, E, device=x.device, dtype=F32)
    dbm = torch.empty(E, device=x.device, dtype=F32)
    _call("mirror_ppeg_bwd", _p(_contig(dy), F32), _p(_contig(x), F32), _p(wm), B, H, E, _p(dx), 0, _p(dwm), _p(dbm), _p(dw7, F32),
          _p(dw5, F32), _p(dw3, F32), _p(db7, F32), _p(db5, F32), _p(db3, F32), launches=3)
    return dx


@_op
def rna_attn_fwd(qkv):
    B, E3 = qkv.shape
    out = torch.empty(B, E3 // 3, device=qkv.device, dtype=F32)
    _call("mirror_rna_attn_fwd", _p(_contig(qkv), F32), B, E3 // 3, None, _p(out))
    return out


@_op
def rna_attn_bwd(qkv, dout):
    B, E3 = qkv.shape
    dqkv = torch.empty_like(qkv)
    _call("mirror_rna_attn_bwd", _p(_contig(qkv), F32), _p(_contig(dout), F32), B, E3 // 3, None, _p(dqkv))
    return dqkv


def _bhrd(t, name):
    """(ptr, ld, head stride, batch stride) of a [B,h,rows,d] bf16 view with unit stride on d."""
    _cuda(t, BF16)
    if t.dim() != 4 or t.stride(3) != 1:
        raise ValueError(f"flash: {name} must be a [B,h,rows,d] view with unit stride on d")
    return t.data_ptr(), t.stride(2), t.stride(1), t.stride(0)


@_op
def flash_softmax_pv(x, y, v, alpha, out, res=None, want_lse=True):
    """out[B,h,R,d] (bf16 view, written in place) = softmax_rows(alpha x y^T) v (+ res) with the probabilities only in
    TMEM / shared memory; returns lse2 [B,h,R] f32 (base-2 log-sum-exp of the scaled logits) for the backward."""
    B, h, R, d = x.shape
    C = y.shape[2]
    if y.shape != (B, h, C, d) or v.shape != (B, h, C, d) or out.shape != (B, h, R, d) or (res is not None and res.shape != out.shape):
        raise ValueError("flash_softmax_pv: shape mismatch")
    a = _lib.FlashArgs()
    a.x, a.x_ld, a.x_hs, a.x_bs = _bhrd(x, "x")
    a.y, a.y_ld, a.y_hs, a.y_bs = _bhrd(y, "y")
    a.v, a.v_ld, a.v_hs, a.v_bs = _bhrd(v, "v")
    a.out, a.o_ld, a.o_hs, a.o_bs = _bhrd(out, "out")
    if res is not None:
        a.res, a.r_ld, a.r_hs, a.r_bs = _bhrd(res, "res")
    a.R, a.C, a.d, a.heads, a.batch, a.alpha = R, C, d, h, B, alpha
    lse2 = torch.empty(B, h, R, device=x.device, dtype=F32) if want_lse else None
    a.lse2 = lse2.data_ptr() if want_lse else None
    _lib.check(_lib.fn("mirror_flash_softmax_pv")(ctypes.byref(a), _stream()), "flash_softmax_pv")
    LAUNCHES[0] += 1
    return lse2


def _flash_out(o, spec, B, h, T, d):
    """spec = (tensor [B,h,T,d] bf16|f32 view, residual bf16 [B,h,ceil(T/row_div),d] view | None, row_div, rscale)"""
    t, res, row_div, rscale = spec
    _cuda(t)
    if t.shape != (B, h, T, d) or t.stride(3) != 1 or t.dtype not in (BF16, F32):
        raise ValueError("flash_bwd: outputs must be [B,h,T,d] bf16 / f32 views with unit stride on d")
    o.ptr, o.is_f32, o.ld, o.hs, o.bs = t.data_ptr(), int(t.dtype == F32), t.stride(2), t.stride(1), t.stride(0)
    o.row_div, o.rscale = row_div, rscale
    if res is not None:
        if res.shape != (B, h, (T + row_div - 1) // row_div, d):
            raise ValueError("flash_bwd: residual must be [B,h,ceil(T/row_div),d]")
        o.res, o.r_ld, o.r_hs, o.r_bs = _bhrd(res, "residual")


@_op
def flash_bwd(a, b, c, dd, alpha, lse2, dot, cols, out1, out2=None):
    """Backward of flash_softmax_pv by recomputation (see mirror_flash_bwd).  a, c: [B,h,T,d]; b, dd: [B,h,L,d] bf16 views;
    lse2 / dot: [B,h,rows] f32 contiguous (rows = T for cols=False, L for cols=True); out1 / out2: (tensor, residual, row_div, rscale)."""
    B, h, T, d = a.shape
    L = b.shape[2]
    if c.shape != a.shape or b.shape != (B, h, L, d) or dd.shape != b.shape:
        raise ValueError("flash_bwd: shape mismatch")
    n_rows = L if cols else T
    if lse2.shape != (B, h, n_rows) or dot.shape != (B, h, n_rows) or not lse2.is_contiguous() or not dot.is_contiguous():
        raise ValueError("flash_bwd: lse2 / dot must be contiguous [B,h,rows]")
    g = _lib.FlashBwdArgs()
    g.a, g.a_ld, g.a_hs, g.a_bs = _bhrd(a, "a")
    g.b, g.b_ld, g.b_hs, g.b_bs = _bhrd(b, "b")
    g.c, g.c_ld, g.c_hs, g.c_bs = _bhrd(c, "c")
    g.dd, g.d_ld, g.d_hs, g.d_bs = _bhrd(dd, "dd")
    g.T, g.L, g.d, g.heads, g.batch, g.cols, g.alpha = T, L, d, h, B, int(cols), alpha
    g.lse2, g.dot = _p(lse2, F32), _p(dot, F32)
    _flash_out(g.out1, out1, B, h, T, d)
    if cols:
        _flash_out(g.out2, out2, B, h, T, d)
    _lib.check(_lib.fn("mirror_flash_bwd")(ctypes.byref(g), _stream()), "flash_bwd")
    LAUNCHES[0] += 1


@_op
def contrastive_stats(x, y, scale, diag0=0):
    """x: [Br,K], y: [Bc,K] bf16 (row stride multiple of 8) -> (lse [Br], diag [Br]) of L = scale * x y^T without L in HBM."""
    Br, Kd = x.shape
    Bc = y.shape[0]
    assert y.shape[1] == Kd and x.stride(1) == 1 and y.stride(1) == 1
    ns = int(_lib.fn("mirror_contrastive_nsplit")(Br, Bc))
    part = torch.empty(ns, Br, 2, device=x.device, dtype=F32)
    lse = torch.empty(Br, device=x.device, dtype=F32)
    diag = torch.empty(Br, device=x.device, dtype=F32)
    _call("mirror_contrastive_stats", _p(x, BF16), x.stride(0), _p(y, BF16), y.stride(0), Br, Bc, Kd, _p(scale, F32), diag0, _p(part), ns,
          _p(lse), _p(diag), launches=2)
    return lse, diag


@_op
def contrastive_grad(x, y, D, E, precise, lo_off, scale, diag0, lse_r, lse_c, a_r, a_c, dscale):
    """-> dx [Br,E] f32 = G y_value (G recomputed tile by tile, see mirror_contrastive_grad); dscale (0-d) is accumulated."""
    Br, Kd = x.shape
    Bc = y.shape[0]
    dx = torch.empty(Br, E, device=x.device, dtype=F32)
    _call("mirror_contrastive_grad", _p(x, BF16), x.stride(0), _p(y, BF16), y.stride(0), Br, Bc, Kd, D, E, int(precise), lo_off,
          _p(scale, F32), diag0, _p(_contig(lse_r), F32), _p(_contig(lse_c), F32), _p(_contig(a_r), F32),
          _p(_contig(a_c), F32) if a_c is not None else None, _p(dx), E, _p(dscale, F32) if dscale is not None else None)
    return dx


@_op
def contrastive_loss(lse_r, lse_c, diag, w_r, w_c, mult, per_sample):
    """-> per-sample losses [B] (per_sample=True) or the 0-d reduction mult * sum_i loss_i."""
    B = lse_r.shape[0]
    out = torch.empty(B if per_sample else (), device=lse_r.device, dtype=F32)
    _call("mirror_contrastive_loss", _p(lse_r, F32), _p(lse_c, F32) if lse_c is not None else None, _p(diag, F32), B, w_r, w_c, mult,
          _p(out) if per_sample else None, None if per_sample else _p(out))
    return out


@_op
def contrastive_coef(g, B, w_r, w_c, mult):
    """upstream gradient (0-d or [B]) -> (a_r, a_c | None): coefficients of the softmax terms in G."""
    g = _contig(g)
    a_r = torch.empty(B, device=g.device, dtype=F32)
    a_c = torch.empty(B, device=g.device, dtype=F32) if w_c != 0.0 else None
    _call("mirror_contrastive_coef", _p(g, F32), 0 if g.dim() == 0 or g.numel() == 1 and B != 1 else 1, B, w_r, w_c, mult, _p(a_r), _p(a_c))
    return a_r, a_c


def _bt_view(t, E):
    """[B,T,E] view with unit stride on E and row stride E -> (ptr tensor, batch stride)."""
    assert t.dim() == 3 and t.stride(2) == 1 and (t.stride(1) == E or t.shape[1] == 1)
    return t.stride(0)


@_op
def masked_mse_fwd(a, b, mask):
    """a, b: [B,T,E] f32 views (row stride E, free batch stride); mask [B,T] -> (loss 0-d, scratch[2])."""
    B, T, E = a.shape
    scratch = torch.empty(2, device=a.device, dtype=F32)
    out = torch.empty((), device=a.device, dtype=F32)
    _call("mirror_masked_mse_fwd", _p(a, F32), _bt_view(a, E), _p(b, F32), _bt_view(b, E), _p(_contig(mask), F32), B, T, E, _p(scratch),
          _p(out), launches=2)
    return out, scratch


@_op
def masked_mse_bwd(a, b, mask, scratch, gout, gw, da, acc_a, db, acc_b):
    B, T, E = a.shape
    _call("mirror_masked_mse_bwd", _p(a, F32), _bt_view(a, E), _p(b, F32), _bt_view(b, E), _p(mask, F32), B, T, E, _p(scratch), _p(gout, F32),
          gw, _p(da, F32), _bt_view(da, E) if da is not None else 0, int(acc_a), _p(db, F32), _bt_view(db, E) if db is not None else 0,
          int(acc_b))


@_op
def gauss_kl_fwd(mu, logvar, B):
    out = torch.empty((), device=mu.device, dtype=F32)
    _call("mirror_gauss_kl_fwd", _p(_contig(mu), F32), _p(_contig(logvar), F32), mu.numel(), B, _p(out))
    return out


@_op
def gauss_kl_bwd_(mu, logvar, B, gout, gw, dmu, dlogvar):
    _call("mirror_gauss_kl_bwd", _p(mu, F32), _p(logvar, F32), mu.numel(), B, _p(gout, F32), gw, _p(_contig(dmu), F32), _p(_contig(dlogvar), F32))


@_op
def sym_kl_fwd(scores, B):
    out = torch.empty((), device=scores.device, dtype=F32)
    _call("mirror_sym_kl_fwd", _p(_contig(scores), F32), B, scores.shape[1], _p(out))
    return out


@_op
def sym_kl_bwd(scores, B, gout, gw):
    d = torch.empty_like(scores)
    _call("mirror_sym_kl_bwd", _p(scores, F32), B, scores.shape[1], _p(gout, F32), gw, _p(d), None)
    return d


@_op
def loss_combine(terms5, weights):
    total = torch.empty((), device=terms5.device, dtype=F32)
    w = (ctypes.c_float * 5)(*weights)
    _call("mirror_loss_combine", _p(_contig(terms5), F32), ctypes.cast(w, ctypes.c_void_p), _p(total))
    return total


# ---- step tail (csrc/optim.cu) and graph-safe dropout -------------------------------------------------------------------
def set_dropout_epoch(counter):
    """Install (or, with None, remove) the int64 DEVICE counter that every dropout launch mixes into its seed (graph replay)."""
    if counter is not None:
        _cuda(counter, torch.int64)
    _lib.check(_lib.fn("mirror_set_dropout_epoch")(counter.data_ptr() if counter is not None else None), "set_dropout_epoch")


@_op
def adam_step_(p, g, m, v, lr, beta1, beta2, eps, weight_decay, decoupled, step, grad_scale=None):
    """In-place Adam / AdamW update of the flat fp32 buffers; lr, step and grad_scale are 0-d / 1-element device tensors."""
    n = p.numel()
    assert g.numel() == n and m.numel() == n and v.numel() == n
    _call("mirror_adam_step", _p(_contig(p), F32), _p(_contig(g), F32), _p(_contig(m), F32), _p(_contig(v), F32), n, _p(lr, F32), beta1, beta2,
          eps, weight_decay, int(decoupled), _p(step, F32), _p(grad_scale, F32) if grad_scale is not None else None)


@_op
def grad_sumsq(g):
    out = torch.empty((), device=g.device, dtype=F32)
    _call("mirror_grad_sumsq", _p(_contig(g), F32), g.numel(), _p(out), launches=1)
    return out


@_op
def tail_scalars_(sumsq=None, max_norm=0.0, coef=None, step=None, clamp_param=None, lo=0.0, hi=0.0):
    """One launch: coef <- clip coefficient, step += 1, clamp_param <- clamp(clamp_param, lo, hi) (each optional)."""
    _call("mirror_tail_scalars", _p(sumsq, F32) if sumsq is not None else None, max_norm, _p(coef, F32) if coef is not None else None,
          _p(step, F32) if step is not None else None, _p(clamp_param, F32) if clamp_param is not None else None, lo, hi)


# ---- input pipeline (mirror_b200/data.py) ----------------------------------------------------------------------------------
@_op
def gather_rows(src, idx, out=None):
    """out[r, :] = float(src[idx[r], :]); src: [n_src, cols] f32 or bf16 (row stride >= cols), idx: int64 [...] -> out [..., cols] f32."""
    if src.dim() != 2 or src.stride(1) != 1 or src.dtype not in (F32, BF16):
        raise ValueError("gather_rows needs a [rows, cols] fp32 / bf16 source with unit column stride")
    _cuda(src), _cuda(idx, torch.int64)
    cols = src.shape[1]
    out = torch.empty(*idx.shape, cols, device=src.device, dtype=F32) if out is None else out
    _call("mirror_gather_rows", src.data_ptr(), int(src.dtype == BF16), src.stride(0), src.shape[0], _p(_contig(idx)), idx.numel(), cols,
          _p(_contig(out), F32))
    return out
