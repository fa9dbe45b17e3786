"""TEST INFRASTRUCTURE: a torch (CPU) re-statement of every entry point of ``mirror_b200.kernels``.

Installed with ``use()`` by the CPU tests so that the host logic — shapes, strides, views, the hand-written
backward passes in ``mirror_b200/ops.py``, the module wiring in ``mirror_b200/models`` and ``losses`` — can be
checked against the oracle without a GPU.  It mimics the kernels' precision plan (bf16 operand storage, fp32
accumulation) so CPU runs also predict the numerical error of the device path.  Never imported by product code.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from mirror_b200 import kernels as K

F32, BF16 = torch.float32, torch.bfloat16


_SAVED = {}


def use():
    """Replace every entry point of mirror_b200.kernels that the emulator re-states (monkeypatched from the test side: the
    product package has no hook for this)."""
    if _SAVED:
        return
    emu = Emu()
    for name in dir(emu):
        if name.startswith("_") or not callable(getattr(emu, name)) or not hasattr(K, name):
            continue
        _SAVED[name] = getattr(K, name)
        setattr(K, name, getattr(emu, name))


def release():
    for name, fn in _SAVED.items():
        setattr(K, name, fn)
    _SAVED.clear()


def hash_u01(seed, idx):
    """mb::hash_u01 (csrc/common.cuh) on numpy uint32."""
    M = np.uint64(0xFFFFFFFF)
    idx = idx.astype(np.uint64)
    seed = int(seed)
    with np.errstate(over="ignore"):
        h = ((idx & M) + np.uint64((seed & 0xFFFFFFFF) * 0x9E3779B1 & 0xFFFFFFFF) + (((idx >> np.uint64(32)) & M) * np.uint64(0x85EBCA77) & M)) & M
        h = h.astype(np.uint32)
        h ^= np.uint32((seed >> 32) & 0xFFFFFFFF)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def keep_mask(seed, shape, p):
    """inverted-dropout multiplier over a dense index space of `shape`."""
    n = int(np.prod(shape))
    u = hash_u01(seed, np.arange(n, dtype=np.uint64))
    return torch.from_numpy(np.where(u >= np.float32(p), np.float32(1.0 / (1.0 - p)), np.float32(0.0))).view(*shape)


def _act(v, act):
    if act == K.ACT_RELU:
        return torch.relu(v)
    if act == K.ACT_GELU:
        return F.gelu(v)
    return v


class Emu:
    launches = 0

    # ------------------------------------------------------------------ GEMM
    def gemm(self, a, b, *, out_f32=None, out_bf16=None, alpha=1.0, bias=None, act=0, drop_p=0.0, drop_seed=0, res=None,
             gamma=1.0, beta=0.0, split_k=1, diag=0.0, more=None, res2=None, gamma2=1.0, res_row_div=1, mode=0, stats=None):
        assert a.dtype == BF16 and b.dtype == BF16
        assert a.stride(-1) == 1 or a.stride(-2) == 1
        assert b.stride(-1) == 1 or b.stride(-2) == 1
        for t in (a, b):  # TMA alignment rules of the real kernel
            ld = t.stride(-2) if t.stride(-1) == 1 else t.stride(-1)
            assert ld % 8 == 0, f"operand ld {ld} not a multiple of 8"
            assert (t.storage_offset() * 2) % 16 == 0, "operand base not 16-byte aligned"
            for s, n in zip(t.stride()[:-2], t.shape[:-2]):
                assert n == 1 or s % 8 == 0, f"batch stride {s}"
        acc = torch.matmul(a.float(), b.float().transpose(-1, -2))
        for (at, bt) in (more or ()):
            assert at.dtype == BF16 and bt.dtype == BF16
            acc = acc + torch.matmul(at.float(), bt.float().transpose(-1, -2))
        if mode != 0:
            return self._gemm_softmax(acc, mode, stats, alpha, res, out_f32, out_bf16)
        if split_k > 1:
            assert out_f32 is not None and out_bf16 is None and bias is None and res is None and act == 0 and drop_p == 0
            out_f32 += alpha * acc.view(out_f32.shape)
            return
        v = alpha * acc
        if diag != 0.0:
            v = v + diag * torch.eye(v.shape[-2], v.shape[-1])
        if bias is not None:
            v = v + bias
        v = _act(v, act)
        if drop_p > 0:
            v = v * keep_mask(drop_seed, tuple(v.shape), drop_p)
        if res is not None:
            assert res.stride(-1) == 1
            rr = res.float()
            if res_row_div > 1:
                rr = rr.repeat_interleave(res_row_div, dim=-2)[..., : v.shape[-2], :]
            assert rr.shape == v.shape
            v = v + gamma * rr
            if res2 is not None:
                assert res2.dtype == BF16 and res2.shape == res.shape and res2.stride() == res.stride()
                r2 = res2.float()
                if res_row_div > 1:
                    r2 = r2.repeat_interleave(res_row_div, dim=-2)[..., : v.shape[-2], :]
                v = v + gamma2 * r2
        if out_f32 is not None:
            assert out_f32.shape == v.shape and out_f32.stride(-1) == 1
            if beta != 0.0:
                v = v + beta * out_f32
            out_f32.copy_(v)
        if out_bf16 is not None:
            assert out_bf16.shape == v.shape and out_bf16.stride(-1) == 1
            out_bf16.copy_(v.to(BF16))

    def gemm_nparts(self, n):
        bn = 256 if (n % 256 == 0 or n >= 1024) else (192 if n % 192 == 0 else 128)
        return 2 * ((n + bn - 1) // bn)

    def _gemm_softmax(self, acc, mode, stats, alpha, res, out_f32, out_bf16):
        """Fused row-softmax epilogue modes: the partials go through `stats` split per half tile as the kernel does."""
        n = acc.shape[-1]
        if mode == 5:
            assert n % 32 == 0 and stats.shape == acc.shape[:-1] and res.dtype == BF16
            v = alpha * res.float() * (acc - stats[..., None])
            out_bf16.copy_(v.to(BF16))
            return
        assert n % 32 == 0 and stats.shape == (*acc.shape[:-1], self.gemm_nparts(n), 2)
        bn = 256 if (n % 256 == 0 or n >= 1024) else (192 if n % 192 == 0 else 128)
        hw = bn // 2
        log2e = 1.4426950408889634
        if mode == 1:
            x2 = alpha * log2e * acc
            stats[..., 0].fill_(float("-inf"))
            stats[..., 1].zero_()
            for p in range(stats.shape[-2]):
                blk = x2[..., p * hw:(p + 1) * hw]
                if blk.shape[-1] == 0:
                    continue
                m = blk.amax(-1)
                stats[..., p, 0] = m
                stats[..., p, 1] = torch.exp2(blk - m[..., None]).sum(-1)
            return
        if mode == 3:
            stats.zero_()
            prod = acc * res.float()
            for p in range(stats.shape[-2]):
                stats[..., p, 0] = prod[..., p * hw:(p + 1) * hw].sum(-1)
            return
        if mode == 2:
            m = stats[..., 0].amax(-1)
            ssum = (stats[..., 1] * torch.exp2(stats[..., 0] - m[..., None])).sum(-1)
            v = torch.exp2(alpha * log2e * acc - m[..., None]) / ssum[..., None]
        else:
            assert mode == 4 and res.dtype == BF16
            v = alpha * res.float() * (acc - stats[..., 0].sum(-1)[..., None])
        if out_f32 is not None:
            out_f32.copy_(v)
        if out_bf16 is not None:
            out_bf16.copy_(v.to(BF16))

    # ------------------------------------------------------------ elementwise
    def cast_bf16(self, src, cols_out=None):
        cols = src.shape[-1]
        cols_out = cols_out or cols
        dst = torch.zeros(*src.shape[:-1], cols_out, dtype=BF16)
        dst[..., :cols] = src.to(BF16)
        return dst

    def cast_split3(self, src, rows_out, cols_out, stack_rows, order):
        rows, cols = src.shape
        hi = torch.zeros(rows_out, cols_out, dtype=BF16)
        lo = torch.zeros(rows_out, cols_out, dtype=BF16)
        hi[:rows, :cols] = src.to(BF16)
        lo[:rows, :cols] = (src - hi[:rows, :cols].float()).to(BF16)
        return torch.cat([hi, lo, hi] if order == 0 else [hi, hi, lo], 0 if stack_rows else 1)

    def copy_rows_(self, src, dst):
        dst.copy_(src)
        return dst

    def axpy_(self, dst, src, alpha=1.0):
        dst += alpha * src
        return dst

    def rowdot(self, a16, b16, sub16=None):
        b = b16.float() - (sub16.float() if sub16 is not None else 0.0)
        return (a16.float() * b).sum(-1)

    def token_fanout_bwd(self, d_full, d_cls, d_tok, B, T, E, device):
        out = torch.zeros(B, T, E) if d_full is None else d_full.clone()
        if d_cls is not None:
            out[:, 0] += d_cls
        if d_tok is not None:
            out[:, 1:] += d_tok
        return out

    def act_fwd(self, pre, act, drop_p=0.0, seed=0, want_bf16=True, want_f32=False):
        v = _act(pre, act)
        if drop_p > 0:
            v = v * keep_mask(seed, tuple(v.shape), drop_p)
        return (v.to(BF16) if want_bf16 else None), (v if want_f32 else None)

    def act_bwd(self, dy, pre, act, drop_p=0.0, seed=0, out16=None, out32=None):
        g = dy.clone()
        if drop_p > 0:
            g = g * keep_mask(seed, tuple(g.shape), drop_p)
        if act == K.ACT_RELU:
            g = g * (pre > 0)
        elif act == K.ACT_GELU:
            x = pre
            cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
            pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
            g = g * (cdf + x * pdf)
        if out16 is not None:
            out16.copy_(g.to(BF16))
        if out32 is not None:
            out32.copy_(g)

    def wsi_assemble_fwd(self, h, cls, N, add):
        h[:, 0, :] = cls
        if add:
            h[:, 1 + N:, :] = h[:, 1:1 + add, :]

    def wsi_embed_bwd(self, dh, h, N, add, dcls):
        g = dh[:, 1:1 + N, :].clone()
        if add:
            g[:, :add, :] += dh[:, 1 + N:, :]
        dcls += dh[:, 0, :].sum(0)
        return (g * (h[:, 1:1 + N, :] > 0)).to(BF16)

    def rank_mask(self, noise, keep):
        rank = torch.argsort(torch.argsort(noise, dim=1, stable=True), dim=1)
        return (rank >= keep).float()

    def mask_pos_fwd_(self, r, mask, tok, tok_stride, pos, first):
        B, T, E = r.shape
        m = torch.zeros(B, T, 1, dtype=torch.bool)
        m[:, first:, 0] = mask != 0
        tokv = tok.reshape(-1)[:E] if tok_stride else tok.reshape(-1)[:1].expand(E)
        r.copy_(torch.where(m, tokv.view(1, 1, E).expand(B, T, E), r) + pos.view(1, T, E))
        return r

    def mask_pos_bwd(self, dy, mask, dtok, tok_stride, dpos, first):
        B, T, E = dy.shape
        dpos += dy.sum(0)
        m = torch.zeros(B, T, 1, dtype=torch.bool)
        m[:, first:, 0] = mask != 0
        g = (dy * m).sum((0, 1))
        if tok_stride:
            dtok[:E] += g
        else:
            dtok[:1] += g.sum()
        return dy * (~m)

    def landmark_fwd(self, qkv, m, seg):
        B, n, E3 = qkv.shape
        E = E3 // 3
        return qkv[:, :, :2 * E].float().view(B, m, seg, 2 * E).sum(2).div(seg).to(BF16)

    def colsum_(self, x, out):
        out += x.float().sum(0)
        return out

    def reparam_fwd(self, mu, logvar, eps):
        return mu + torch.exp(0.5 * logvar) * eps

    def reparam_bwd_(self, dz, logvar, eps, dmu, dlogvar):
        dmu += dz
        dlogvar += dz * eps * 0.5 * torch.exp(0.5 * logvar)

    # ------------------------------------------------------- norms / softmax
    def layernorm_fwd(self, x, gamma, beta, eps, n_out=None, pad=0, want_bf16=True, want_f32=False, rows=None):
        B, X, E = x.shape
        S = rows or X
        x = x[:, :S]
        n_out = n_out or S
        mean = x.mean(-1)
        var = x.var(-1, unbiased=False)
        rstd = torch.rsqrt(var + eps)
        y = (x - mean[..., None]) * rstd[..., None] * gamma + beta
        full = torch.zeros(B, n_out, E)
        full[:, pad:pad + S] = y
        return (full.to(BF16) if want_bf16 else None), (full if want_f32 else None), mean, rstd

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, pad, dx, add, dgamma, dbeta):
        B, X, E = x.shape
        S = mean.shape[1]
        g = dy[:, pad:pad + S].float()
        xh = (x[:, :S] - mean[..., None]) * rstd[..., None]
        gg = g * gamma
        d = torch.zeros(B, X, E)
        d[:, :S] = rstd[..., None] * (gg - gg.mean(-1, keepdim=True) - xh * (gg * xh).mean(-1, keepdim=True))
        dgamma += (g * xh).sum((0, 1))
        dbeta += g.sum((0, 1))
        dx.copy_(d + add if add is not None else d)

    def softmax_fwd(self, x, want_bf16=True, want_f32=False):
        p = torch.softmax(x, -1)
        return (p.to(BF16) if want_bf16 else None), (p if want_f32 else None)

    def softmax_bwd(self, y16, dy, scale=1.0, want_bf16=True, want_f32=False):
        y = y16.float()
        d = scale * y * (dy - (dy * y).sum(-1, keepdim=True))
        return (d.to(BF16) if want_bf16 else None), (d if want_f32 else None)

    def l2norm_fwd(self, x, eps):
        norm = x.norm(dim=-1).clamp_min(eps)
        return x / norm[:, None], norm

    def l2norm_bwd(self, dy, x, norm):
        y = x / norm[:, None]
        return (dy - y * (y * dy).sum(-1, keepdim=True)) / norm[:, None]

    # --------------------------------------------------------------- nystrom
    def res_conv_fwd(self, qkv, w):
        B, n, E3 = qkv.shape
        E = E3 // 3
        h = w.shape[0]
        v = qkv[:, :, 2 * E:].float().view(B, n, h, E // h).permute(0, 2, 1, 3)
        o = F.conv2d(v, w.view(h, 1, -1, 1), padding=(w.shape[1] // 2, 0), groups=h)
        return o.permute(0, 2, 1, 3).reshape(B, n, E).to(BF16)

    @torch.enable_grad()
    def res_conv_bwd(self, dout16, qkv, w, dw):
        B, n, E3 = qkv.shape
        E = E3 // 3
        h = w.shape[0]
        v = qkv[:, :, 2 * E:].float().view(B, n, h, E // h).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
        ww = w.view(h, 1, -1, 1).clone().requires_grad_(True)
        o = F.conv2d(v, ww, padding=(w.shape[1] // 2, 0), groups=h)
        go = dout16.float().view(B, n, h, E // h).permute(0, 2, 1, 3)
        gv, gw = torch.autograd.grad(o, (v, ww), go)
        dw += gw.view(h, -1)
        return gv.permute(0, 2, 1, 3).reshape(B, n, E).to(BF16)

    def pinv_init(self, a2, z16=None, scratch=None):
        ax = a2.abs()
        rs, cs = ax.sum(-1), ax.sum(-2)
        scratch = torch.zeros(8) if scratch is None else scratch
        scratch[0], scratch[1] = rs.max(), cs.max()
        scratch[3], scratch[4] = float(rs.flatten().argmax()), float(cs.flatten().argmax())
        z = a2.transpose(-1, -2) / (scratch[0] * scratch[1])
        z = z.contiguous().to(BF16)
        if z16 is not None:
            z16.copy_(z)
            z = z16
        return z, scratch

    def pinv_init_bwd(self, gz0, z0_16, scratch, gx, accumulate):
        c, r = scratch[0], scratch[1]
        D = c * r
        dD = -(gz0 * z0_16.float()).sum() / D
        m = gz0.shape[-1]
        v = gz0.transpose(-1, -2) / D
        v = v.contiguous()
        vf = v.view(-1, m, m)
        ir, ic = int(scratch[3]), int(scratch[4])
        vf[ir // m, ir % m, :] += dD * r
        vf[ic // m, :, ic % m] += dD * c
        if accumulate:
            gx += v
        else:
            gx.copy_(v)

    def pinv_init_softmax_bwd(self, ga2, gz0, z0_16, a2_16, scratch, scale):
        g = ga2.clone()
        self.pinv_init_bwd(gz0, z0_16, scratch, g, True)
        return self.softmax_bwd(a2_16, g, scale)[0]

    # ------------------------------------------------------------------ ppeg
    def ppeg_fwd(self, x, w7, w5, w3, b7, b5, b3, H):
        B, S, E = x.shape
        wm = w7.view(E, 7, 7).clone()
        wm[:, 1:6, 1:6] += w5.view(E, 5, 5)
        wm[:, 2:5, 2:5] += w3.view(E, 3, 3)
        wm[:, 3, 3] += 1.0
        f = x[:, 1:].transpose(1, 2).reshape(B, E, H, H)
        y = F.conv2d(f, wm.view(E, 1, 7, 7), b7 + b5 + b3, padding=3, groups=E)
        return torch.cat([x[:, :1], y.flatten(2).transpose(1, 2)], 1), wm.view(E, 49).t().contiguous()

    @torch.enable_grad()
    def ppeg_bwd(self, dy, x, wm, H, dw7, dw5, dw3, db7, db5, db3):
        B, S, E = x.shape
        w = wm.t().reshape(E, 1, 7, 7).clone().requires_grad_(True)
        f = x[:, 1:].transpose(1, 2).reshape(B, E, H, H).clone().requires_grad_(True)
        y = F.conv2d(f, w, None, padding=3, groups=E)
        g = dy[:, 1:].transpose(1, 2).reshape(B, E, H, H)
        gf, gw = torch.autograd.grad(y, (f, w), g)
        gw = gw.view(E, 7, 7)
        dw7 += gw.reshape(E, 49)
        dw5 += gw[:, 1:6, 1:6].reshape(E, 25)
        dw3 += gw[:, 2:5, 2:5].reshape(E, 9)
        gb = g.sum((0, 2, 3))
        db7 += gb
        db5 += gb
        db3 += gb
        return torch.cat([dy[:, :1], gf.flatten(2).transpose(1, 2)], 1)

    # -------------------------------------------------------------- rna attn
    def _rna(self, qkv):
        B, E3 = qkv.shape
        E = E3 // 3
        t = qkv.view(B, 3, 12, E // 12)
        q, k, v = t[:, 0], t[:, 1], t[:, 2]
        a = torch.softmax(q @ k.transpose(-1, -2) * (E // 12) ** -0.5, -1)
        return (a @ v).transpose(1, 2).reshape(B, E)

    def rna_attn_fwd(self, qkv):
        return self._rna(qkv)

    @torch.enable_grad()
    def rna_attn_bwd(self, qkv, dout):
        q = qkv.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(self._rna(q), q, dout)
        return g

    # ---------------------------------------------------------------- losses
    def flash_softmax_pv(self, x, y, v, alpha, out, res=None, want_lse=True):
        """unnormalised probabilities 2^(a2 (S - max)): fp32 row sum, bf16 operands of the value product, O / l in fp32 (the kernel's order)"""
        for t in (x, y, v, out) + ((res,) if res is not None else ()):
            assert t.dtype == BF16 and t.stride(3) == 1 and all(s % 8 == 0 for s in t.stride()[:3])
        a2 = alpha * 1.4426950408889634
        s_ = torch.matmul(x.float(), y.float().transpose(-1, -2))
        mx = s_.max(-1, keepdim=True).values
        p = torch.exp2(a2 * (s_ - mx))
        l = p.sum(-1, keepdim=True)
        o = torch.matmul(p.to(BF16).float(), v.float()) / l
        if res is not None:
            o = o + res.float()
        out.copy_(o.to(BF16))
        return (a2 * mx + torch.log2(l)).squeeze(-1) if want_lse else None

    def flash_bwd(self, a, b, c, dd, alpha, lse2, dot, cols, out1, out2=None):
        a2 = alpha * 1.4426950408889634
        s_ = torch.matmul(a.float(), b.float().transpose(-1, -2))   # [.., T, L]
        dp = torch.matmul(c.float(), dd.float().transpose(-1, -2))
        if cols:   # statistics run along the block dim (softmax rows = L)
            p_ = torch.exp2(a2 * s_ - lse2[..., None, :])
            ds = alpha * p_ * (dp - dot[..., None, :])
        else:
            p_ = torch.exp2(a2 * s_ - lse2[..., :, None])
            ds = alpha * p_ * (dp - dot[..., :, None])

        def store(spec, val):
            t, res, row_div, rscale = spec
            if res is not None:
                val = val + rscale * res.float().repeat_interleave(row_div, dim=-2)[..., : val.shape[-2], :]
            t.copy_(val.to(t.dtype))

        store(out1, torch.matmul(ds.to(BF16).float(), b.float()))
        if cols:
            store(out2, torch.matmul(p_.to(BF16).float(), dd.float()))

    def contrastive_stats(self, x, y, scale, diag0=0):
        assert x.dtype == BF16 and y.dtype == BF16 and x.stride(0) % 8 == 0 and y.stride(0) % 8 == 0
        l = scale * (x.float() @ y.float().T)
        i = torch.arange(x.shape[0])
        return torch.logsumexp(l, 1), l[i, i + diag0]

    def contrastive_grad(self, x, y, D, E, precise, lo_off, scale, diag0, lse_r, lse_c, a_r, a_c, dscale):
        Br, Kd = x.shape
        assert D % 64 == 0 and E <= D and (Kd == 3 * D if precise else Kd == D)
        raw = x.float() @ y.float().T
        l = scale * raw
        pr = a_r[:, None] * torch.exp(l - lse_r[:, None])
        g = pr.clone()
        i = torch.arange(Br)
        if a_c is not None:
            g = g + a_c[None, :] * torch.exp(l - lse_c[None, :])
            g[i, i + diag0] -= a_c[i + diag0]
        g[i, i + diag0] -= a_r
        if dscale is not None:
            d = pr.clone()
            d[i, i + diag0] -= a_r
            dscale += (d * raw).sum()
        g = g * scale
        yf = y.float()
        if precise:  # G = hi + lo (both bf16), value = hi + lo blocks of y; the lo*lo term is dropped like in the kernel
            gh = g.to(BF16).float()
            gl = (g - gh).to(BF16).float()
            yh, yl = yf[:, :D], yf[:, lo_off:lo_off + D]
            dx = gh @ yh + gl @ yh + gh @ yl
        else:
            dx = g.to(BF16).float() @ yf[:, :D]
        return dx[:, :E].contiguous()

    def contrastive_loss(self, lse_r, lse_c, diag, w_r, w_c, mult, per_sample):
        l = w_r * (lse_r - diag)
        if w_c != 0.0:
            l = l + w_c * (lse_c - diag)
        return l if per_sample else l.sum() * mult

    def contrastive_coef(self, g, B, w_r, w_c, mult):
        v = (g.reshape(-1) * mult).expand(B) if g.numel() == 1 else g * mult
        return (v * w_r).contiguous(), ((v * w_c).contiguous() if w_c != 0.0 else None)

    def masked_mse_fwd(self, a, b, mask):
        E = a.shape[-1]
        num = (((a - b) ** 2).sum(-1) / E * mask).sum()
        den = mask.sum()
        return num / den, torch.stack([num, den])

    def masked_mse_bwd(self, a, b, mask, scratch, gout, gw, da, acc_a, db, acc_b):
        E = a.shape[-1]
        v = gout * gw * 2.0 / (E * scratch[1]) * mask[..., None] * (a - b)
        if da is not None:
            da.copy_(da + v if acc_a else v)
        if db is not None:
            db.copy_(db - v if acc_b else -v)

    def gauss_kl_fwd(self, mu, logvar, B):
        return 0.5 / B * (logvar.exp() + mu ** 2 - 1 - logvar).sum()

    def gauss_kl_bwd_(self, mu, logvar, B, gout, gw, dmu, dlogvar):
        k = gout * gw / B
        dmu += k * mu
        dlogvar += k * 0.5 * (logvar.exp() - 1)

    def _symkl(self, scores, B):
        lw, lr = F.log_softmax(scores[:B], -1), F.log_softmax(scores[B:], -1)
        return 0.5 / B * ((lr.exp() - lw.exp()) * (lr - lw)).sum()

    def sym_kl_fwd(self, scores, B):
        return self._symkl(scores, B)

    @torch.enable_grad()
    def sym_kl_bwd(self, scores, B, gout, gw):
        s = scores.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(self._symkl(s, B), s)
        return g * gout * gw

    def loss_combine(self, terms5, weights):
        return (terms5 * torch.tensor(weights, dtype=F32)).sum()

    def gather_rows(self, src, idx, out=None):
        y = src[idx.clamp(0, src.shape[0] - 1)].float()
        if out is not None:
            out.copy_(y)
            return out
        return y

    # ---- step tail (csrc/optim.cu)
    def set_dropout_epoch(self, counter):
        self.drop_epoch = counter  # the emulated dropout ignores it (graph replay does not exist on the CPU)

    def adam_step_(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, decoupled, step, grad_scale=None):
        lr, t = float(lr), float(step)
        gg = g * (float(grad_scale) if grad_scale is not None else 1.0)
        if weight_decay != 0.0:
            if decoupled:
                p.mul_(1.0 - lr * weight_decay)
            else:
                gg = gg + weight_decay * p
        m.mul_(beta1).add_(gg, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
        p.sub_((lr / (1 - beta1 ** t)) * m / (v.sqrt() / (1 - beta2 ** t) ** 0.5 + eps))

    def grad_sumsq(self, g):
        return (g.double() ** 2).sum().float()

    def tail_scalars_(self, sumsq=None, max_norm=0.0, coef=None, step=None, clamp_param=None, lo=0.0, hi=0.0):
        if coef is not None:
            coef.fill_(min(1.0, max_norm / (float(sumsq) ** 0.5 + 1e-6)) if sumsq is not None and max_norm > 0 else 1.0)
        if step is not None:
            step.add_(1.0)
        if clamp_param is not None:
            clamp_param.clamp_(lo, hi)
