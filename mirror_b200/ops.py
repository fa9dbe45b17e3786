"""Autograd functions of the MIRROR hot path, written on top of the C-ABI kernels.

Forward AND backward of every function are explicit sequences of kernel
launches (``mirror_b200.kernels``); PyTorch autograd only chains them.
Precision plan (SURVEY.md §7 "Precision"): bf16 tensor-core operands, fp32
accumulation, fp32 residual stream / LayerNorm / softmax statistics / head
outputs / losses.
"""
import math
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K

F32, BF16 = torch.float32, torch.bfloat16
WSI_HEADS = 8  # models/mirror.py:302
PINV_ITERS = 6  # models/mirror.py:304

_seed_state = {"base": None, "n": 0}


def next_seed() -> int:
    """Host-side counter-based seed for the hashed dropout masks (no device sync).  The stream restarts whenever
    torch.manual_seed() changes torch's seed (resume, per-rank seeding after the first forward)."""
    base = torch.initial_seed() & 0xFFFFFFFF
    if _seed_state["base"] != base:
        _seed_state["base"], _seed_state["n"] = base, 0
    _seed_state["n"] += 1
    return ((base * 0x9E3779B1) ^ (_seed_state["n"] * 0x85EBCA77)) & 0x7FFFFFFFFFFFFFFF


class Side:
    """Holder for a bf16 side copy handed to a Function: custom_fwd(cast_inputs=float32) converts every floating-point
    TENSOR argument under torch.autocast -- a bf16 copy passed bare would come out as fp32 (and be rejected by the GEMM).
    A plain object is passed through untouched."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _r8(x):
    return (x + 7) // 8 * 8


def _cfwd(fn):
    return torch.amp.custom_fwd(fn, device_type="cuda", cast_inputs=torch.float32)


def _cbwd(fn):
    return torch.amp.custom_bwd(fn, device_type="cuda")


def _T(x):
    return x.transpose(-1, -2)


def _split_k(M, N, Kdim):
    tiles = ((M + 127) // 128) * ((N + 255) // 256 if (N % 256 == 0 or N >= 1024) else (N + 127) // 128)
    kb = (Kdim + 63) // 64
    return max(1, min(148 // max(tiles, 1), kb // 4 if kb >= 8 else 1, 32))


def wgrad(dy16, x16, out_rows, out_cols):
    """dW[out_rows,out_cols] (f32) = dy16^T @ x16 with the contraction over rows, split-K over the SMs.
    dy16: [rows, >=out_rows] bf16 view, x16: [rows, >=out_cols] bf16 view (row strides multiples of 8)."""
    rows = dy16.shape[0]
    dw = torch.zeros(out_rows, out_cols, device=dy16.device, dtype=F32)
    K.gemm(_T(dy16[:, :out_rows]), _T(x16[:, :out_cols]), out_f32=dw, split_k=_split_k(out_rows, out_cols, rows))
    return dw


def colsum(x, cols):
    out = torch.zeros(cols, device=x.device, dtype=F32)
    K.colsum_(x[:, :cols], out)
    return out


# ----------------------------------------------------------------------------------------------
PRECISE_ROWS = 1024  # Linears with at most this many rows are weight-bandwidth-bound: run them fp32-grade (split-3 bf16)


class LinearFn(Function):
    """y = [res +] dropout(x @ W^T + b) [+ row_res broadcast over rows].  nn.Linear of the heads, the RNA encoder and the
    style MLPs (models/mirror.py:70,74,470-495,594-605,823-827; timm Mlp fc1/fc2) with the following Dropout and residual
    add of Block.forward (:149-152) fused into the GEMM epilogue.  x16: optional bf16 copy of x.

    rows <= PRECISE_ROWS (everything whose M is the batch size): operands are split into bf16 hi/lo parts and the three
    cross products run as ONE tensor-core GEMM over a tripled contraction, forward and backward, so these layers are
    fp32-accurate; their cost is the weight stream, not the FLOPs.  Larger inputs use plain bf16 operands."""

    @staticmethod
    @_cfwd
    def forward(ctx, x, weight, bias, row_res, x16, res, drop_p, seed):
        N, Kd = weight.shape
        lead = x.shape[:-1]
        rows = x.numel() // Kd
        x16 = x16.t if isinstance(x16, Side) else x16
        if x16 is not None and (x16.dtype != BF16 or x16.shape != x.shape):
            x16 = None
        Kp = _r8(Kd)
        precise = rows <= PRECISE_ROWS
        y = torch.empty(*lead, N, device=x.device, dtype=F32)  # returned as-is (no view) so callers may update it in place
        assert row_res is None or res is None
        r = None
        if row_res is not None:
            r = row_res.reshape(1, N).expand(rows, N)
        elif res is not None:
            r = res.contiguous().view(rows, N)
        if precise:
            x2 = x if x.dim() == 2 and x.stride(1) == 1 else x.contiguous().view(rows, Kd)
            K.gemm(K.cast_split3(x2, rows, Kp, False, 0), K.cast_split3(weight, N, Kp, False, 1), out_f32=y.view(rows, N),
                   bias=bias, res=r, drop_p=drop_p, drop_seed=seed)
            ctx.save_for_backward(x2, weight)
        else:
            if x16 is None or Kp != Kd:
                x2 = x.reshape(rows, Kd) if x.is_contiguous() or x.dim() == 2 else x.contiguous().view(rows, Kd)
                x16 = K.cast_bf16(x2, Kp)
            else:
                x16 = x16.reshape(rows, Kd)
            w16 = K.cast_bf16(weight, Kp)
            K.gemm(x16[:, :Kd], w16[:, :Kd], out_f32=y.view(rows, N), bias=bias, res=r, drop_p=drop_p, drop_seed=seed)
            ctx.save_for_backward(x16, w16)
        ctx.meta = (rows, N, Kd, bias is not None, tuple(row_res.shape) if row_res is not None else None, res is not None,
                    tuple(x.shape), drop_p, seed, precise)
        return y

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        xs, ws = ctx.saved_tensors
        rows, N, Kd, has_b, has_rr, has_res, xshape, drop_p, seed, precise = ctx.meta
        Np, Kp = _r8(N), _r8(Kd)
        dy2 = dy.reshape(rows, N)
        if dy2.stride(1) != 1 or (dy2.stride(0) != N and rows > 1):
            dy2 = dy2.contiguous()
        dx = dw = db = drr = dres = None
        if precise:
            g = dy2
            if drop_p > 0:
                g = torch.empty(rows, N, device=dy.device, dtype=F32)
                K.act_bwd(dy2.view(1, rows, N), None, K.ACT_NONE, drop_p, seed, out32=g.view(1, rows, N))
            Rp = _r8(rows)
            if ctx.needs_input_grad[0]:
                dx = torch.empty(rows, Kd, device=dy.device, dtype=F32)
                wst = K.cast_split3(ws, Np, Kp, True, 1)                                  # [3Np, Kp] = (hi; hi; lo)
                K.gemm(K.cast_split3(g, rows, Np, False, 0), _T(wst[:, :Kd]), out_f32=dx)  # dX = dY W
                dx = dx.view(xshape)
            if ctx.needs_input_grad[1]:
                dw = torch.empty(N, Kd, device=dy.device, dtype=F32)
                gst = K.cast_split3(g, Rp, Np, True, 0)                                    # [3Rp, Np]
                xst = K.cast_split3(xs, Rp, Kp, True, 1)                                   # [3Rp, Kp]
                K.gemm(_T(gst[:, :N]), _T(xst[:, :Kd]), out_f32=dw)                        # dW = dY^T X
            if has_b and ctx.needs_input_grad[2]:
                db = colsum(g, N)
        else:
            if drop_p > 0:
                dy16 = torch.empty(rows, Np, device=dy.device, dtype=BF16)
                K.act_bwd(dy2.view(1, rows, N), None, K.ACT_NONE, drop_p, seed, out16=dy16[:, :N].unsqueeze(0))
            else:
                dy16 = K.cast_bf16(dy2, Np)
            if ctx.needs_input_grad[0]:
                dx = torch.empty(rows, Kd, device=dy.device, dtype=F32)
                K.gemm(dy16[:, :N], _T(ws[:, :Kd]), out_f32=dx)
                dx = dx.view(xshape)
            if ctx.needs_input_grad[1]:
                dw = wgrad(dy16, xs, N, Kd)
            if has_b and ctx.needs_input_grad[2]:
                db = colsum(dy16, N)
        if has_rr is not None and ctx.needs_input_grad[3]:
            drr = colsum(dy2, N).view(has_rr)  # row_res is added after the dropout
        if has_res and ctx.needs_input_grad[5]:
            dres = dy
        return dx, dw, db, drr, None, dres, None, None


def linear(x, weight, bias=None, row_res=None, x16=None, res=None, drop_p=0.0):
    return LinearFn.apply(x, weight, bias, row_res, Side(x16) if x16 is not None else None, res, drop_p,
                          next_seed() if drop_p > 0 else 0)


class StackRowsFn(Function):
    """[a; b] for two [B,E] f32 row blocks (a may be a strided view such as the cls rows)."""

    @staticmethod
    @_cfwd
    def forward(ctx, a, b):
        B, E = a.shape
        out = torch.empty(2 * B, E, device=a.device, dtype=F32)
        K.copy_rows_(a, out[:B])
        K.copy_rows_(b, out[B:])
        ctx.B = B
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return dy[:ctx.B], dy[ctx.B:]


def stack_rows(a, b):
    return StackRowsFn.apply(a, b)


# ----------------------------------------------------------------------------------------------
class LayerNormFn(Function):
    """nn.LayerNorm over the last dim; returns (y f32, y bf16 copy [non-differentiable]).  ``keep``: x is [B,X,E] and only
    the first ``keep`` tokens of every slide are normalised and returned ([B,keep,E], contiguous) -- LayerNorm followed by
    the reference's ``h[:, :-add_length]`` (models/mirror.py:372) without materialising the dropped rows."""

    @staticmethod
    @_cfwd
    def forward(ctx, x, weight, bias, eps, keep=None):
        E = x.shape[-1]
        ctx.shape = tuple(x.shape)
        ctx.set_materialize_grads(False)  # the bf16 copy never has a gradient: no zero-filled stand-in for it
        if keep is not None and keep < x.shape[1]:
            x3 = x.contiguous()
            out_shape = (x.shape[0], keep, E)
        else:
            keep = None
            x3 = x.contiguous().view(1, -1, E)
            out_shape = ctx.shape
        y16, y32, mean, rstd = K.layernorm_fwd(x3, weight, bias, eps, want_bf16=True, want_f32=True, rows=keep)
        ctx.save_for_backward(x3, weight, mean, rstd)
        y16 = y16.view(out_shape)
        ctx.mark_non_differentiable(y16)
        return y32.view(out_shape), y16

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy, _unused):
        if dy is None:
            return None, None, None, None, None
        x3, weight, mean, rstd = ctx.saved_tensors
        E = x3.shape[-1]
        dx = torch.empty_like(x3)
        dg = torch.zeros(E, device=dy.device, dtype=F32)
        db = torch.zeros(E, device=dy.device, dtype=F32)
        K.layernorm_bwd(dy.contiguous().view(x3.shape[0], -1, E), x3, weight, mean, rstd, 0, dx, None, dg, db)
        return dx.view(ctx.shape), dg, db, None, None


class TokenFanoutFn(Function):
    """h [B,T,E] -> (cls = h[:,0,:] (copy), h, h[:,1:,:] (view)).  The three ways the model reads its encoder output
    (models/mirror.py:889-905); the backward merges the three gradients in one pass instead of autograd's
    zeros + slice-copy + add chain over the full token matrix."""

    @staticmethod
    @_cfwd
    def forward(ctx, h):
        h = h.contiguous()
        ctx.shape = tuple(h.shape)
        ctx.set_materialize_grads(False)  # unused outputs must not cost a zero-filled token matrix
        cls = torch.empty(h.shape[0], h.shape[2], device=h.device, dtype=F32)
        K.copy_rows_(h[:, 0, :], cls)
        return cls, h.view_as(h), h[:, 1:, :]

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, d_cls, d_full, d_tok):
        B, T, E = ctx.shape
        if d_cls is None and d_full is None and d_tok is None:
            return None
        dev = next(t for t in (d_cls, d_full, d_tok) if t is not None).device
        if d_tok is not None and d_tok.stride(2) != 1:
            d_tok = d_tok.contiguous()
        return K.token_fanout_bwd(d_full, d_cls, d_tok, B, T, E, dev)


def token_fanout(h):
    return TokenFanoutFn.apply(h)


def drop_first_token(r):
    """r[:, 1:, :] whose backward writes the padded gradient in one pass (no zeros + slice copy)."""
    return TokenFanoutFn.apply(r)[2]


def layer_norm(x, weight, bias, eps, keep=None):
    return LayerNormFn.apply(x, weight, bias, eps, keep)


# ----------------------------------------------------------------------------------------------
class ActFn(Function):
    """dropout(act(x)) for the small RNA / style MLP activations (exact-erf GELU)."""

    @staticmethod
    @_cfwd
    def forward(ctx, pre, act, drop_p, seed):
        pre = pre.contiguous()
        _, y = K.act_fwd(pre, act, drop_p, seed, want_bf16=False, want_f32=True)
        ctx.save_for_backward(pre)
        ctx.cfg = (act, drop_p, seed)
        return y

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        (pre,) = ctx.saved_tensors
        act, drop_p, seed = ctx.cfg
        dx = torch.empty_like(pre)
        C = pre.shape[-1]
        K.act_bwd(dy.contiguous().view(1, -1, C), pre.view(1, -1, C), act, drop_p, seed, out32=dx.view(1, -1, C))
        return dx, None, None, None


def activation(pre, act, drop_p=0.0):
    return ActFn.apply(pre, act, drop_p, next_seed() if drop_p > 0 else 0)


# ----------------------------------------------------------------------------------------------
class L2NormFn(Function):
    """F.normalize(x, dim=-1, eps) on a [rows, E] view (rows may be strided, e.g. the cls rows)."""

    @staticmethod
    @_cfwd
    def forward(ctx, x, eps):
        y, norm = K.l2norm_fwd(x, eps)
        ctx.save_for_backward(x, norm)
        return y

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        x, norm = ctx.saved_tensors
        return K.l2norm_bwd(dy.contiguous(), x, norm), None


def l2_normalize(x, eps):
    return L2NormFn.apply(x, eps)


# ----------------------------------------------------------------------------------------------
class WsiEmbedFn(Function):
    """_fc1 (Linear+ReLU) + wrap-around square padding + cls token (models/mirror.py:652-665) -> h [B,S,E] f32."""

    @staticmethod
    @_cfwd
    def forward(ctx, wsi, w, b, cls):
        B, N, Dw = wsi.shape
        E = w.shape[0]
        H = int(math.ceil(math.sqrt(N)))
        add = H * H - N
        S = H * H + 1
        Dp = _r8(Dw)
        # split-3 bf16 operands ([hi|lo|hi] x [hi|hi|lo]): an fp32-grade product, so the ReLU mask matches the reference
        x16 = K.cast_split3(wsi.contiguous().view(B * N, Dw), B * N, Dp, False, 0).view(B, N, 3 * Dp)
        w16 = K.cast_split3(w.contiguous(), E, Dp, False, 1)
        h = torch.empty(B, S, E, device=wsi.device, dtype=F32)
        K.gemm(x16, w16.unsqueeze(0).expand(B, E, 3 * Dp), out_f32=h[:, 1:1 + N, :], bias=b, act=K.ACT_RELU)
        K.wsi_assemble_fwd(h, cls.reshape(E), N, add)
        ctx.save_for_backward(x16, h)
        ctx.meta = (B, N, Dw, E, add)
        return h

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dh):
        x16, h = ctx.saved_tensors
        B, N, Dw, E, add = ctx.meta
        dcls = torch.zeros(E, device=dh.device, dtype=F32)
        dpre = K.wsi_embed_bwd(dh.contiguous(), h, N, add, dcls)
        dw = wgrad(dpre.view(B * N, E), x16.view(B * N, -1), E, Dw)  # the "hi" block of x16 is its first Dw columns
        db = colsum(dpre.view(B * N, E), E)
        return None, dw, db, dcls.view(1, 1, E)


# ----------------------------------------------------------------------------------------------
def _heads(t, col0, E, h=WSI_HEADS):
    """[B, rows, W] -> view [B, h, rows, d] of columns col0 .. col0+E."""
    B, rows, _ = t.shape
    return t[:, :, col0:col0 + E].unflatten(-1, (h, E // h)).permute(0, 2, 1, 3)


_AB_F32_DXN = os.environ.get("MIRROR_B200_AB_F32_DXN") == "1"                    # A/B switches (measurement only)
_FUSED_PINV_BWD = os.environ.get("MIRROR_B200_FUSED_PINV_BWD") == "1"  # opt-in: measured 1.30 ms vs 1.01 ms for the two kernels
_AB_NO_FLASH = os.environ.get("MIRROR_B200_AB_NO_FLASH") == "1"  # A/B switch (measurement only): materialised attn1 / attn3
_AB_NO_DOTS = os.environ.get("MIRROR_B200_AB_NO_DOTS") == "1"  # A/B switch (measurement only): two-pass softmax backward everywhere


def _softmax_gemm(a, b, rows, cols, alpha, want_f32=False):
    """softmax_rows(alpha * a @ b^T) -> (bf16, f32 | None) without the logits touching HBM: two GEMM passes over the
    (K = head_dim) product, row statistics then normalised probabilities (MIRROR_GEMM_ROWSTATS / SOFTMAX)."""
    batch, dev = a.shape[:-2], a.device
    if cols % 32:  # unfused: fp32 logits + row softmax kernel
        s_ = torch.empty(*batch, rows, cols, device=dev, dtype=F32)
        K.gemm(a, b, out_f32=s_, alpha=alpha)
        return K.softmax_fwd(s_, want_f32=want_f32)
    st = K.softmax_stats(batch, rows, cols, dev)
    K.gemm(a, b, alpha=alpha, mode=K.GEMM_ROWSTATS, stats=st)
    p16 = torch.empty(*batch, rows, cols, device=dev, dtype=BF16)
    p32 = torch.empty(*batch, rows, cols, device=dev, dtype=F32) if want_f32 else None
    K.gemm(a, b, alpha=alpha, mode=K.GEMM_SOFTMAX, stats=st, out_bf16=p16, out_f32=p32)
    return p16, p32


def _softmax_bwd_gemm(a, b, p16, alpha, dots=None):
    """d logits = alpha * P * (G - rowsum(G * P)) with G = a @ b^T never materialised (MIRROR_GEMM_ROWDOT / SOFTMAX_BWD).
    ``dots``: rowsum(G * P) when the caller has it cheaper (P fed O = P V, so rowsum(G * P) = dO . O row by row)."""
    batch, dev = a.shape[:-2], a.device
    rows, cols = p16.shape[-2:]
    if cols % 32:
        g = torch.empty(*batch, rows, cols, device=dev, dtype=F32)
        K.gemm(a, b, out_f32=g)
        return K.softmax_bwd(p16, g, alpha)[0]
    ds = torch.empty_like(p16)
    if dots is not None:
        K.gemm(a, b, alpha=alpha, mode=K.GEMM_SOFTMAX_BWD_DOT, stats=dots, res=p16, out_bf16=ds)
        return ds
    st = K.softmax_stats(batch, rows, cols, dev)
    K.gemm(a, b, mode=K.GEMM_ROWDOT, stats=st, res=p16)
    K.gemm(a, b, alpha=alpha, mode=K.GEMM_SOFTMAX_BWD, stats=st, res=p16, out_bf16=ds)
    return ds


class NystromLayerFn(Function):
    """TransLayer: x + NystromAttention(LayerNorm(x)) (models/mirror.py:295-314; nystrom_attention forward).

    Differences from the reference evaluation order, all exact in real arithmetic: q's 1/sqrt(d) is folded into the three
    similarity GEMMs; (attn1 z)(attn3 v) is evaluated as attn1 (z (attn3 v)); each Moore-Penrose step
    z' = 0.25 z (13I - xz(15I - xz(7I - xz))) is evaluated in the residual E = I - a2 z as z' = z + z E (I + E + 0.25 E^2),
    the same cubic without the cancellation between O(10)-sized terms that costs bf16 operands their accuracy.
    """

    @staticmethod
    @_cfwd
    def forward(ctx, h, ln_w, ln_b, qkv_w, out_w, out_b, conv_w, drop_p, seed, eps=1e-5, per_slide_scale=False):
        B, S, E = h.shape
        hd = WSI_HEADS
        d, m = E // hd, E // 2
        pad = (m - S % m) % m
        n = S + pad
        seg = math.ceil(S / m)
        scale = d ** -0.5
        dev = h.device
        h = h.contiguous()
        xn16, _, mean, rstd = K.layernorm_fwd(h, ln_w, ln_b, eps, n_out=n, pad=pad)
        wqkv16 = K.cast_bf16(qkv_w)
        qkv = torch.empty(B, n, 3 * E, device=dev, dtype=BF16)
        K.gemm(xn16.view(B * n, E), wqkv16, out_bf16=qkv.view(B * n, 3 * E))
        lm = K.landmark_fwd(qkv, m, seg)
        q, k, v = _heads(qkv, 0, E), _heads(qkv, E, E), _heads(qkv, 2 * E, E)
        ql, kl = _heads(lm, 0, E), _heads(lm, E, E)
        flash = d % 8 == 0 and d <= 128 and not _AB_NO_FLASH  # fused softmax products (csrc/flash_nystrom.cu): attn1 / attn3 never reach HBM
        a1 = a3 = lse1 = lse3 = None
        kv = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
        if flash:
            lse3 = K.flash_softmax_pv(ql, k, v, scale, kv)                        # kv = softmax(ql k^T) v
        else:
            a1, _ = _softmax_gemm(q, kl, n, m, scale)
            a3, _ = _softmax_gemm(ql, k, m, n, scale)
            K.gemm(a3, _T(v), out_bf16=kv)
        a2_16, a2_32 = _softmax_gemm(ql, kl, m, m, scale, want_f32=True)
        if per_slide_scale and B > 1:  # variable-length bags: every slide is its own "batch" (the reference at B = 1 per slide)
            z16 = torch.empty_like(a2_32, dtype=BF16)
            scratch = torch.empty(B, 8, device=dev, dtype=F32)
            for b in range(B):
                K.pinv_init(a2_32[b], z16[b], scratch[b])
        else:
            z16, scratch = K.pinv_init(a2_32)
        iters = []
        mm = (B, hd, m, m)
        for _ in range(PINV_ITERS):
            # z' = 0.25 z (13 I - xz (15 I - xz (7 I - xz))) written in the residual E = I - xz:
            #   z' = z + z F,  F = E + E G1,  G1 = E + 0.25 E^2      (= z + z (E + E^2 + 0.25 E^3))
            # -- the same cubic, but no cancellation between O(10) terms; the leading terms are added in the fp32 epilogue.
            Em = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(a2_16, _T(z16), out_bf16=Em, alpha=-1.0, diag=1.0)              # E  = I - a2 z
            G1 = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(Em, _T(Em), out_bf16=G1, alpha=0.25, res=Em)                    # G1 = E + 0.25 E E
            Fm = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(Em, _T(G1), out_bf16=Fm, res=Em)                                # F  = E + E G1
            zn16 = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(z16, _T(Fm), out_bf16=zn16, res=z16)                            # z' = z + z F  (bf16 iterate: measured
            iters += [z16, Em, G1, Fm]                                             #  +3e-4 grad rel-L2 vs an fp32 master copy)
            z16 = zn16
        w_ = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
        K.gemm(z16, _T(kv), out_bf16=w_)
        rc = K.res_conv_fwd(qkv, conv_w.reshape(hd, -1))
        o16 = torch.empty(B, n, E, device=dev, dtype=BF16)
        if flash:
            lse1 = K.flash_softmax_pv(q, kl, w_, scale, _heads(o16, 0, E), res=_heads(rc, 0, E))   # out = softmax(q kl^T) w + res_conv(v)
        else:
            K.gemm(a1, _T(w_), out_bf16=_heads(o16, 0, E), res=_heads(rc, 0, E))
        wout16 = K.cast_bf16(out_w)
        y = torch.empty(B, S, E, device=dev, dtype=F32)
        K.gemm(o16[:, pad:, :], wout16.unsqueeze(0).expand(B, E, E), out_f32=y, bias=out_b, drop_p=drop_p, drop_seed=seed, res=h)
        ctx.save_for_backward(h, ln_w, mean, rstd, xn16, wqkv16, qkv, lm, lse1 if flash else a1, a2_16, lse3 if flash else a3, scratch,
                              z16, kv, w_, o16, wout16, conv_w, rc, *iters)
        ctx.meta = (B, S, E, pad, n, seg, scale, drop_p, seed, flash)
        ctx.per_slide_scale = bool(per_slide_scale and B > 1)
        return y

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        (h, ln_w, mean, rstd, xn16, wqkv16, qkv, lm, a1, a2_16, a3, scratch, zf16, kv, w_, o16, wout16,
         conv_w, rc, *iters) = ctx.saved_tensors
        B, S, E, pad, n, seg, scale, drop_p, seed, flash = ctx.meta
        hd = WSI_HEADS
        d, m = E // hd, E // 2
        dev = dy.device
        dy = dy.contiguous()
        if flash:
            return NystromLayerFn._backward_flash(ctx, dy)
        q, k, v = _heads(qkv, 0, E), _heads(qkv, E, E), _heads(qkv, 2 * E, E)
        ql, kl = _heads(lm, 0, E), _heads(lm, E, E)
        mm = (B, hd, m, m)

        # ---- to_out: y = h + drop(o16[pad:] @ Wout^T + b)
        dyd = torch.empty(B, n, E, device=dev, dtype=BF16)
        if pad:
            dyd[:, :pad, :].zero_()  # the zero rows the sequence was front-padded with
        K.act_bwd(dy, None, K.ACT_NONE, drop_p, seed, out16=dyd[:, pad:, :])
        do16 = torch.empty(B, n, E, device=dev, dtype=BF16)
        K.gemm(dyd.view(B * n, E), _T(wout16), out_bf16=do16.view(B * n, E))
        d_out_w = wgrad(dyd.view(B * n, E), o16.view(B * n, E), E, E)
        d_out_b = colsum(dyd.view(B * n, E), E)
        del dyd
        do_h = _heads(do16, 0, E)

        # ---- out = a1 @ w + res_conv(v)
        # da1 = dO w^T stays in TMEM; its row dots with a1 are dO . (a1 w) = dO . (out - res_conv(v)), per token and head,
        # so the ROWDOT pass over the n x m product is not needed (the 226 MB value residual is kept for this)
        dots1 = None
        if not _AB_NO_DOTS:
            dots1 = K.rowdot(do16.view(B, n, hd, d), o16.view(B, n, hd, d), rc.view(B, n, hd, d)).permute(0, 2, 1).contiguous()
        ds1 = _softmax_bwd_gemm(do_h, w_, a1, scale, dots=dots1)
        dw16 = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
        K.gemm(_T(a1), _T(do_h), out_bf16=dw16)

        # ---- w = z @ kv ; kv = a3 @ v
        gz16 = torch.empty(mm, device=dev, dtype=BF16)
        K.gemm(dw16, kv, out_bf16=gz16)
        dkv16 = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
        K.gemm(_T(zf16), _T(dw16), out_bf16=dkv16)
        # da3 = dkv v^T; its row dots with a3 are dkv . kv (kv = a3 v), so one pass suffices
        ds3 = _softmax_bwd_gemm(dkv16, v, a3, scale, dots=None if _AB_NO_DOTS else K.rowdot(dkv16, kv))
        d_conv = torch.zeros(hd, conv_w.numel() // hd, device=dev, dtype=F32)
        dvc = K.res_conv_bwd(do16, qkv, conv_w.reshape(hd, -1), d_conv)              # conv^T(dO), bf16 [B,n,E]
        del do16
        dqkv16 = torch.empty(B, n, 3 * E, device=dev, dtype=BF16)
        K.gemm(_T(a3), _T(dkv16), out_bf16=_heads(dqkv16, 2 * E, E), res=_heads(dvc, 0, E))  # dv = a3^T dkv + conv^T(dO)
        del dvc

        # ---- Moore-Penrose iterations, reversed.  With gEn := -g_E every sum of products is ONE multi-term GEMM:
        #   gF   = z^T g                gG1q = 0.25 E^T gF
        #   gEn  = -(gF G1^T + gG1q E^T + E^T gG1q) - gF - 4 gG1q
        #   g   <- g + g F^T + a2^T gEn           and after the loop   g_a2 = sum_it gEn_it z_it^T
        gens = []
        for it in reversed(range(PINV_ITERS)):
            z16, Em, G1, Fm = iters[4 * it:4 * it + 4]
            gF = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(_T(z16), _T(gz16), out_bf16=gF)
            gG1q = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(_T(Em), _T(gF), out_bf16=gG1q, alpha=0.25)
            gEn = torch.empty(mm, device=dev, dtype=BF16)
            K.gemm(gF, G1, more=[(gG1q, Em), (_T(Em), _T(gG1q))], out_bf16=gEn, alpha=-1.0, res=gF, gamma=-1.0, res2=gG1q,
                   gamma2=-4.0)
            gzn16 = torch.empty(mm, device=dev, dtype=BF16)
            gzn32 = torch.empty(mm, device=dev, dtype=F32) if it == 0 else None  # only d/dz0 is needed in fp32
            K.gemm(gz16, Fm, more=[(_T(a2_16), _T(gEn))], out_f32=gzn32, out_bf16=gzn16, res=gz16)
            gens.append((gEn, z16))
            gz32, gz16 = gzn32, gzn16
        ga2 = torch.empty(mm, device=dev, dtype=F32)
        K.gemm(gens[0][0], gens[0][1], more=gens[1:], out_f32=ga2)
        del gens
        if ctx.per_slide_scale:
            for b in range(B):
                K.pinv_init_bwd(gz32[b], iters[0][b], scratch[b], ga2[b], True)
            ds2, _ = K.softmax_bwd(a2_16, ga2, scale)
        elif _FUSED_PINV_BWD and m % 32 == 0 and m <= 512:  # iters[0] = z0 (bf16)
            ds2 = K.pinv_init_softmax_bwd(ga2, gz32, iters[0], a2_16, scratch, scale)
        else:
            K.pinv_init_bwd(gz32, iters[0], scratch, ga2, True)
            ds2, _ = K.softmax_bwd(a2_16, ga2, scale)
        del ga2, gz32, gz16

        # ---- similarities (1/sqrt(d) already folded into ds*).  Landmark gradients first (each a two-term GEMM), then dq / dk
        # with the landmark-mean backward fused as a row-broadcast residual: row t of q receives dql[t // seg] / seg.
        dlm16 = torch.empty(B, m, 2 * E, device=dev, dtype=BF16)  # bf16 like dq / dk themselves: the lean epilogue reads it
        K.gemm(ds2, _T(kl), more=[(ds3, _T(k))], out_bf16=_heads(dlm16, 0, E))             # dql = ds2 kl + ds3 k
        K.gemm(_T(ds1), _T(q), more=[(_T(ds2), _T(ql))], out_bf16=_heads(dlm16, E, E))     # dkl = ds1^T q + ds2^T ql
        K.gemm(ds1, _T(kl), out_bf16=_heads(dqkv16, 0, E), res=_heads(dlm16, 0, E), gamma=1.0 / seg, res_row_div=seg)      # dq
        K.gemm(_T(ds3), _T(ql), out_bf16=_heads(dqkv16, E, E), res=_heads(dlm16, E, E), gamma=1.0 / seg, res_row_div=seg)  # dk
        del ds1, ds2, ds3, dlm16

        # ---- to_qkv and LayerNorm
        dxn = torch.empty(B, n, E, device=dev, dtype=F32 if _AB_F32_DXN else BF16)  # bf16 like dqkv itself: lean epilogue, half the LN-backward read
        K.gemm(dqkv16.view(B * n, 3 * E), _T(wqkv16), **{"out_f32" if _AB_F32_DXN else "out_bf16": dxn.view(B * n, E)})
        d_qkv_w = wgrad(dqkv16.view(B * n, 3 * E), xn16.view(B * n, E), 3 * E, E)
        dg = torch.zeros(E, device=dev, dtype=F32)
        db = torch.zeros(E, device=dev, dtype=F32)
        dh = torch.empty_like(h)
        K.layernorm_bwd(dxn, h, ln_w, mean, rstd, pad, dh, dy, dg, db)
        return dh, dg, db, d_qkv_w, d_out_w, d_out_b, d_conv.view(conv_w.shape), None, None, None, None


def _nystrom_backward_flash(ctx, dy):
    """Backward of NystromLayerFn with the probability matrices recomputed block by block (csrc/flash_nystrom.cu):
    four fused launches (two orientations x attn1 / attn3) replace the softmax-gradient passes and the five products that
    read or wrote the [n x m] matrices."""
    (h, ln_w, mean, rstd, xn16, wqkv16, qkv, lm, lse1, a2_16, lse3, scratch, zf16, kv, w_, o16, wout16,
     conv_w, rc, *iters) = ctx.saved_tensors
    B, S, E, pad, n, seg, scale, drop_p, seed, _ = ctx.meta
    hd = WSI_HEADS
    d, m = E // hd, E // 2
    dev = dy.device
    q, k, v = _heads(qkv, 0, E), _heads(qkv, E, E), _heads(qkv, 2 * E, E)
    ql, kl = _heads(lm, 0, E), _heads(lm, E, E)
    mm = (B, hd, m, m)

    # ---- to_out: y = h + drop(o16[pad:] @ Wout^T + b)
    dyd = torch.empty(B, n, E, device=dev, dtype=BF16)
    if pad:
        dyd[:, :pad, :].zero_()
    K.act_bwd(dy, None, K.ACT_NONE, drop_p, seed, out16=dyd[:, pad:, :])
    do16 = torch.empty(B, n, E, device=dev, dtype=BF16)
    K.gemm(dyd.view(B * n, E), _T(wout16), out_bf16=do16.view(B * n, E))
    d_out_w = wgrad(dyd.view(B * n, E), o16.view(B * n, E), E, E)
    d_out_b = colsum(dyd.view(B * n, E), E)
    del dyd
    do_h = _heads(do16, 0, E)

    # ---- out = attn1 w + res_conv(v): row dots dO . (out - res_conv(v)) per token and head, then the key-stationary pass
    dots1 = K.rowdot(do16.view(B, n, hd, d), o16.view(B, n, hd, d), rc.view(B, n, hd, d)).permute(0, 2, 1).contiguous()
    dlm32 = torch.empty(B, m, 2 * E, device=dev, dtype=F32)   # partial landmark gradients (dql | dkl) accumulated in fp32
    dw16 = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
    K.flash_bwd(kl, q, w_, do_h, scale, lse1, dots1, True, (_heads(dlm32, E, E), None, 1, 1.0), (dw16, None, 1, 1.0))  # ds1^T q, attn1^T dO

    # ---- w = z kv ; kv = attn3 v
    gz16 = torch.empty(mm, device=dev, dtype=BF16)
    K.gemm(dw16, kv, out_bf16=gz16)
    dkv16 = torch.empty(B, hd, m, d, device=dev, dtype=BF16)
    K.gemm(_T(zf16), _T(dw16), out_bf16=dkv16)
    dots3 = K.rowdot(dkv16, kv)
    K.flash_bwd(ql, k, dkv16, v, scale, lse3, dots3, False, (_heads(dlm32, 0, E), None, 1, 1.0))                   # ds3 k
    d_conv = torch.zeros(hd, conv_w.numel() // hd, device=dev, dtype=F32)
    dvc = K.res_conv_bwd(do16, qkv, conv_w.reshape(hd, -1), d_conv)              # conv^T(dO), bf16 [B,n,E]

    # ---- Moore-Penrose iterations, reversed (see NystromLayerFn.backward)
    gens = []
    gz32 = None
    for it in reversed(range(PINV_ITERS)):
        z16, Em, G1, Fm = iters[4 * it:4 * it + 4]
        gF = torch.empty(mm, device=dev, dtype=BF16)
        K.gemm(_T(z16), _T(gz16), out_bf16=gF)
        gG1q = torch.empty(mm, device=dev, dtype=BF16)
        K.gemm(_T(Em), _T(gF), out_bf16=gG1q, alpha=0.25)
        gEn = torch.empty(mm, device=dev, dtype=BF16)
        K.gemm(gF, G1, more=[(gG1q, Em), (_T(Em), _T(gG1q))], out_bf16=gEn, alpha=-1.0, res=gF, gamma=-1.0, res2=gG1q,
               gamma2=-4.0)
        gzn16 = torch.empty(mm, device=dev, dtype=BF16)
        gzn32 = torch.empty(mm, device=dev, dtype=F32) if it == 0 else None
        K.gemm(gz16, Fm, more=[(_T(a2_16), _T(gEn))], out_f32=gzn32, out_bf16=gzn16, res=gz16)
        gens.append((gEn, z16))
        gz32, gz16 = gzn32, gzn16
    ga2 = torch.empty(mm, device=dev, dtype=F32)
    K.gemm(gens[0][0], gens[0][1], more=gens[1:], out_f32=ga2)
    del gens
    if ctx.per_slide_scale:
        for b in range(B):
            K.pinv_init_bwd(gz32[b], iters[0][b], scratch[b], ga2[b], True)
    else:
        K.pinv_init_bwd(gz32, iters[0], scratch, ga2, True)
    ds2, _ = K.softmax_bwd(a2_16, ga2, scale)
    del ga2, gz32, gz16

    # ---- landmark gradients: dql = ds2 kl + (ds3 k), dkl = ds2^T ql + (ds1^T q); then the token-stationary passes, which add the
    # landmark-mean backward (token t receives d_landmark[t // seg] / seg) and the value residual in their epilogues
    dlm16 = torch.empty(B, m, 2 * E, device=dev, dtype=BF16)
    K.gemm(ds2, _T(kl), out_bf16=_heads(dlm16, 0, E), res=_heads(dlm32, 0, E))
    K.gemm(_T(ds2), _T(ql), out_bf16=_heads(dlm16, E, E), res=_heads(dlm32, E, E))
    dqkv16 = torch.empty(B, n, 3 * E, device=dev, dtype=BF16)
    K.flash_bwd(k, ql, v, dkv16, scale, lse3, dots3, True, (_heads(dqkv16, E, E), _heads(dlm16, E, E), seg, 1.0 / seg),
                (_heads(dqkv16, 2 * E, E), _heads(dvc, 0, E), 1, 1.0))                                                  # dk, dv
    K.flash_bwd(q, kl, do_h, w_, scale, lse1, dots1, False, (_heads(dqkv16, 0, E), _heads(dlm16, 0, E), seg, 1.0 / seg))  # dq
    del do16, dvc, dlm16, dlm32, ds2

    # ---- to_qkv and LayerNorm
    dxn = torch.empty(B, n, E, device=dev, dtype=BF16)
    K.gemm(dqkv16.view(B * n, 3 * E), _T(wqkv16), out_bf16=dxn.view(B * n, E))
    d_qkv_w = wgrad(dqkv16.view(B * n, 3 * E), xn16.view(B * n, E), 3 * E, E)
    dg = torch.zeros(E, device=dev, dtype=F32)
    db = torch.zeros(E, device=dev, dtype=F32)
    dh = torch.empty_like(h)
    K.layernorm_bwd(dxn, h, ln_w, mean, rstd, pad, dh, dy, dg, db)
    return dh, dg, db, d_qkv_w, d_out_w, d_out_b, d_conv.view(conv_w.shape), None, None, None, None


NystromLayerFn._backward_flash = staticmethod(_nystrom_backward_flash)


def nystrom_layer(h, norm_w, norm_b, qkv_w, out_w, out_b, conv_w, drop_p, eps=1e-5, per_slide_scale=False):
    return NystromLayerFn.apply(h, norm_w, norm_b, qkv_w, out_w, out_b, conv_w, drop_p, next_seed() if drop_p > 0 else 0, eps,
                                per_slide_scale)


# ----------------------------------------------------------------------------------------------
class PpegFn(Function):
    """PPEG (models/mirror.py:317-331) as one merged 7x7 depthwise stencil over token-major activations."""

    @staticmethod
    @_cfwd
    def forward(ctx, x, w7, b7, w5, b5, w3, b3):
        B, S, E = x.shape
        H = int(round(math.sqrt(S - 1)))
        x = x.contiguous()
        y, wm = K.ppeg_fwd(x, w7.reshape(E, 49), w5.reshape(E, 25), w3.reshape(E, 9), b7, b5, b3, H)
        ctx.save_for_backward(x, wm)
        ctx.H = H
        return y

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        x, wm = ctx.saved_tensors
        E = x.shape[-1]
        z = lambda *s: torch.zeros(*s, device=dy.device, dtype=F32)
        dw7, dw5, dw3, db7, db5, db3 = z(E, 49), z(E, 25), z(E, 9), z(E), z(E), z(E)
        dx = K.ppeg_bwd(dy.contiguous(), x, wm, ctx.H, dw7, dw5, dw3, db7, db5, db3)
        return dx, dw7.view(E, 1, 7, 7), db7, dw5.view(E, 1, 5, 5), db5, dw3.view(E, 1, 3, 3), db3


# ----------------------------------------------------------------------------------------------
class MaskPosFn(Function):
    """random_masking's token replacement + learned position table (models/mirror.py:521-527,549,636-643,692-693).
    r: [B,T,E]; mask: [B,T-first] (1 = masked); tok: [E] or scalar; pos: [T,E].  r is overwritten."""

    @staticmethod
    @_cfwd
    def forward(ctx, r, mask, tok, pos, first):
        B, T, E = r.shape
        tok_stride = 1 if tok.numel() == E and E > 1 else (1 if tok.numel() > 1 else 0)
        K.mask_pos_fwd_(r, mask, tok.reshape(-1), tok_stride, pos.reshape(T, E).contiguous(), first)
        ctx.mark_dirty(r)
        ctx.save_for_backward(mask)
        ctx.meta = (tok_stride, first, tuple(tok.shape), tuple(pos.shape))
        return r

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        tok_stride, first, tshape, pshape = ctx.meta
        B, T, E = dy.shape
        dtok = torch.zeros(tshape, device=dy.device, dtype=F32)
        dpos = torch.zeros(T, E, device=dy.device, dtype=F32)
        dr = K.mask_pos_bwd(dy.contiguous(), mask, dtok.view(-1), tok_stride, dpos, first)
        return dr, None, dtok, dpos.view(pshape), None


# ----------------------------------------------------------------------------------------------
class RnaAttnFn(Function):
    """SDPA over the 12 chunks of one embedding + dim-major interleave (models/mirror.py:77-99)."""

    @staticmethod
    @_cfwd
    def forward(ctx, qkv):
        qkv = qkv.contiguous()
        ctx.save_for_backward(qkv)
        return K.rna_attn_fwd(qkv)

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        return K.rna_attn_bwd(qkv, dout.contiguous())


# ----------------------------------------------------------------------------------------------
class ReparamFn(Function):
    """z = mu + exp(0.5*logvar)*eps (models/mirror.py:830-833 with the N(0,1) draw made explicit)."""

    @staticmethod
    @_cfwd
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        ctx.save_for_backward(logvar, eps)
        return K.reparam_fwd(mu, logvar, eps)

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dmu = torch.zeros_like(logvar)
        dlv = torch.zeros_like(logvar)
        K.reparam_bwd_(dz.contiguous(), logvar, eps, dmu, dlv)
        return dmu, dlv, None


# ----------------------------------------------------------------------------------------------
def _r64(x):
    return (x + 63) // 64 * 64


def contrastive_operands(w, r, precise):
    """bf16 operands of the fused contrastive kernels: plain [B, Dp] copies, or split-3 [B, 3 Dp] (w: hi|lo|hi, r: hi|hi|lo,
    so that w3 . r3 = hi hi + lo hi + hi lo whichever of the two is the row operand).  Dp = E rounded up to 64 (zeros)."""
    B, E = w.shape
    Dp = _r64(E)
    if precise:
        return K.cast_split3(w, B, Dp, False, 0), K.cast_split3(r, r.shape[0], Dp, False, 1), Dp
    return K.cast_bf16(w, Dp), K.cast_bf16(r, Dp), Dp


def _gather_rows(t, group):
    """all-gather of a [B_local, ...] tensor along dim 0 -> [world * B_local, ...] (rank-major).  No gradient flows through it:
    ClipLossFn's backward forms the exact gradient of the local rows from the gathered statistics instead."""
    import torch.distributed as dist
    t = t.contiguous()
    out = torch.empty(dist.get_world_size(group) * t.shape[0], *t.shape[1:], device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t, group=group)
    return out


class ClipLossFn(Function):
    """Contrastive loss over logits = scale * W R^T: ClipLoss (losses/mirror_loss.py:37-52, w_row=w_col=0.5) and
    InfoNCE's implicit-negative branch (losses/info_nce.py:144-164), reduction 'mean' | 'sum' | 'none'.
    `scale` is a 0-d device tensor.  The B x B logits never reach HBM (csrc/contrastive.cu): two statistics passes (row
    log-sum-exp of W R^T and of R W^T) and two gradient passes that recompute the logits tile by tile in TMEM and feed
    G = dLoss/dlogits straight into a second tcgen05.mma (dW = G R, dR = G^T W).
    Up to PRECISE_ROWS samples the operands are split-3 bf16 (fp32-grade products): the temperature multiplies every
    rounding error of the similarities by ~14-100.

    ``group`` (a torch.distributed process group; None = the reference's rank-local negatives): global
    negatives.  The embeddings of all ranks are all-gathered; this rank evaluates its B_local rows of W_loc R_glob^T and of
    R_loc W_glob^T (its rows and its columns of the global logits matrix), returns the mean over ITS samples, and a second,
    tiny all-gather of the two log-sum-exp vectors lets its backward form the exact gradient of the global-batch loss with
    respect to its local embeddings -- contributions of the other ranks' loss terms included -- so no gradient is sent back
    through the gather.  The gradient is that of SUM over ranks of the per-rank losses (every rank receives the same upstream
    gradient), which DDP's averaging turns into the gradient of their mean = the single-process global-batch loss."""

    @staticmethod
    @_cfwd
    def forward(ctx, w, r, scale, w_row, w_col, reduction="mean", group=None):
        B, E = w.shape
        w, r = w.contiguous(), r.contiguous()
        rank = 0
        wg, rg = w, r
        if group is not None:
            import torch.distributed as dist
            rank = dist.get_rank(group)
            wg, rg = _gather_rows(w, group), _gather_rows(r, group)
        Bg = wg.shape[0]
        diag0 = rank * B
        precise = Bg <= PRECISE_ROWS
        w16, r16, Dp = contrastive_operands(wg, rg, precise)   # operands of the whole (global) batch; local rows are a slice
        w16l, r16l = w16[diag0:diag0 + B], r16[diag0:diag0 + B]
        scale = scale.reshape(()).contiguous()
        lse_r, diag = K.contrastive_stats(w16l, r16, scale, diag0)
        lse_c = K.contrastive_stats(r16l, w16, scale, diag0)[0] if w_col != 0.0 else None
        mult = 1.0 / B if reduction == "mean" else 1.0
        loss = K.contrastive_loss(lse_r, lse_c, diag, w_row, w_col, mult, reduction == "none")
        lse_rg, lse_cg = lse_r, lse_c
        if group is not None:  # statistics of every row / column of the global logits, for the backward of the local rows
            lse_rg = _gather_rows(lse_r, group)
            lse_cg = _gather_rows(lse_c, group) if lse_c is not None else None
        ctx.save_for_backward(w16, r16, scale, lse_rg, lse_cg)
        ctx.meta = (B, Bg, diag0, E, Dp, w_row, w_col, precise, mult)
        return loss

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, gout):
        w16, r16, scale, lse_rg, lse_cg = ctx.saved_tensors
        B, Bg, diag0, E, Dp, w_row, w_col, precise, mult = ctx.meta
        w16l, r16l = w16[diag0:diag0 + B], r16[diag0:diag0 + B]
        dscale = torch.zeros((), device=gout.device, dtype=F32)
        a_r, a_c = K.contrastive_coef(gout.contiguous(), B, w_row, w_col, mult)
        if lse_cg is None:  # one-sided: the column terms vanish (a_c = None), the column LSE is never read
            lse_cg = lse_rg
        lse_r, lse_c = lse_rg[diag0:diag0 + B], lse_cg[diag0:diag0 + B]
        a_rg, a_cg = a_r, a_c
        if Bg != B:  # coefficients of the other ranks' samples: the same upstream gradient on every rank (see the class docstring)
            if gout.numel() != 1:
                raise NotImplementedError("global negatives need a scalar reduction ('mean' or 'sum')")
            a_rg = a_r[:1].expand(Bg).contiguous()
            a_cg = a_c[:1].expand(Bg).contiguous() if a_c is not None else None
        # w3 = hi|lo|hi (lo block at Dp), r3 = hi|hi|lo (lo block at 2 Dp)
        dw = K.contrastive_grad(w16l, r16, Dp, E, precise, 2 * Dp, scale, diag0, lse_r, lse_cg, a_r, a_cg, dscale)
        if a_c is None:
            a_c0 = torch.zeros_like(a_r)
            dr = K.contrastive_grad(r16l, w16, Dp, E, precise, Dp, scale, diag0, lse_c, lse_rg, a_c0, a_rg, None)
        else:
            dr = K.contrastive_grad(r16l, w16, Dp, E, precise, Dp, scale, diag0, lse_c, lse_rg, a_c, a_rg, dscale)
        return dw, dr, dscale, None, None, None, None


def clip_loss(w, r, scale, w_row=0.5, w_col=0.5, reduction="mean", group=None):
    return ClipLossFn.apply(w, r, scale, w_row, w_col, reduction, group)


# ----------------------------------------------------------------------------------------------
class MaskedMseFn(Function):
    """sum_rows mask * mean_e (a-b)^2 / sum mask (losses/mirror_loss.py:98-103); a, b: [B,T,E] views."""

    @staticmethod
    @_cfwd
    def forward(ctx, a, b, mask):
        out, scratch = K.masked_mse_fwd(a, b, mask.contiguous())
        ctx.save_for_backward(a, b, mask, scratch)
        return out

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, g):
        a, b, mask, scratch = ctx.saved_tensors
        da = torch.empty(a.shape, device=a.device, dtype=F32) if ctx.needs_input_grad[0] else None
        db = torch.empty(b.shape, device=a.device, dtype=F32) if ctx.needs_input_grad[1] else None
        K.masked_mse_bwd(a, b, mask, scratch, g.contiguous(), 1.0, da, False, db, False)
        return da, db, None


class GaussKlFn(Function):
    """0.5/B * sum (exp(lv) + mu^2 - 1 - lv) over stacked modalities (losses/mirror_loss.py:105-112)."""

    @staticmethod
    @_cfwd
    def forward(ctx, mu, lv, B):
        mu, lv = mu.contiguous(), lv.contiguous()
        ctx.save_for_backward(mu, lv)
        ctx.B = B
        return K.gauss_kl_fwd(mu, lv, B)

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, g):
        mu, lv = ctx.saved_tensors
        dmu, dlv = torch.zeros_like(mu), torch.zeros_like(lv)
        K.gauss_kl_bwd_(mu, lv, ctx.B, g.contiguous(), 1.0, dmu, dlv)
        return dmu, dlv, None


class SymKlFn(Function):
    """0.5*(KL(r||w)+KL(w||r)) batchmean over softmaxed prototype scores (losses/mirror_loss.py:114-119).
    scores: [2B,P], WSI rows first."""

    @staticmethod
    @_cfwd
    def forward(ctx, scores, B):
        scores = scores.contiguous()
        ctx.save_for_backward(scores)
        ctx.B = B
        return K.sym_kl_fwd(scores, B)

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, g):
        (scores,) = ctx.saved_tensors
        return K.sym_kl_bwd(scores, ctx.B, g.contiguous(), 1.0), None


_WEIGHT_VECTORS = {}


class CombineFn(Function):
    """total = sum_i w_i * term_i (losses/mirror_loss.py:121-127)."""

    @staticmethod
    @_cfwd
    def forward(ctx, terms, weights):
        ctx.weights = weights
        return K.loss_combine(terms.contiguous(), weights)

    @staticmethod
    @once_differentiable
    @_cbwd
    def backward(ctx, g):
        # five scalars (bookkeeping, not arithmetic of the hot path).  The weight vector is uploaded ONCE per (weights,
        # device): torch.tensor(list, device=cuda) is a blocking copy -- as the first node of every backward it made the
        # host wait for the whole forward (26 ms per step) and the GPU then idle while the backward was being enqueued.
        key = (tuple(float(w) for w in ctx.weights), g.device)
        w = _WEIGHT_VECTORS.get(key)
        if w is None:
            w = _WEIGHT_VECTORS[key] = torch.tensor(key[0], device=g.device, dtype=F32)
        return g.reshape(1) * w, None


class StackScalarsFn(Function):
    """Pack five 0-d device scalars into one [5] tensor (pure pointer bookkeeping: five 4-byte copies)."""

    @staticmethod
    def forward(ctx, *xs):
        out = torch.empty(len(xs), device=xs[0].device, dtype=F32)
        for i, x in enumerate(xs):
            K.copy_rows_(x.reshape(1, 1), out[i:i + 1].view(1, 1))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return tuple(g[i] for i in range(g.shape[0]))
