"""Stand-in for PyPI ``nystrom_attention~=0.0.14`` (requirements.txt:5).

TEST INFRASTRUCTURE. Restates the published NystromAttention forward used by
``models/mirror.py:299-309``: front zero padding to a multiple of the landmark
count, bias-free qkv, segment-mean landmarks, three row softmaxes, a 6-step
Moore-Penrose iteration whose initial scale is a max over the WHOLE tensor,
a 33-tap depthwise value residual and an output projection with dropout.
Parity of this file itself is unpinned upstream (no reference test exists).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def moore_penrose_iter_pinv(x, iters=6):
    ax = x.abs()
    scale = ax.sum(dim=-1).max() * ax.sum(dim=-2).max()  # global maxima
    z = x.transpose(-1, -2) / scale
    eye = torch.eye(x.shape[-1], device=x.device, dtype=x.dtype).unsqueeze(0)
    for _ in range(iters):
        xz = x @ z
        z = 0.25 * z @ (13 * eye - (xz @ (15 * eye - (xz @ (7 * eye - xz)))))
    return z


class NystromAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256,
                 pinv_iterations=6, residual=True, residual_conv_kernel=33,
                 eps=1e-8, dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        self.eps = eps
        self.heads = heads
        self.num_landmarks = num_landmarks
        self.pinv_iterations = pinv_iterations
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            k = residual_conv_kernel
            self.res_conv = nn.Conv2d(heads, heads, (k, 1), padding=(k // 2, 0),
                                      groups=heads, bias=False)

    def forward(self, x, mask=None, return_attn=False):
        assert mask is None and not return_attn, "shim covers the MIRROR call only"
        b, n, _ = x.shape
        h, m = self.heads, self.num_landmarks
        if n % m:
            x = F.pad(x, (0, 0, m - n % m, 0), value=0.0)
        npad = x.shape[1]
        q, k, v = self.to_qkv(x).chunk(3, dim=-1)
        q, k, v = (t.reshape(b, npad, h, -1).transpose(1, 2) for t in (q, k, v))
        q = q * self.scale
        seg = math.ceil(n / m)
        q_l = q.reshape(b, h, npad // seg, seg, -1).sum(dim=3) / seg
        k_l = k.reshape(b, h, npad // seg, seg, -1).sum(dim=3) / seg
        a1 = (q @ k_l.transpose(-1, -2)).softmax(dim=-1)
        a2 = (q_l @ k_l.transpose(-1, -2)).softmax(dim=-1)
        a3 = (q_l @ k.transpose(-1, -2)).softmax(dim=-1)
        out = (a1 @ moore_penrose_iter_pinv(a2, self.pinv_iterations)) @ (a3 @ v)
        if self.residual:
            out = out + self.res_conv(v)
        out = out.transpose(1, 2).reshape(b, npad, -1)
        return self.to_out(out)[:, -n:]
