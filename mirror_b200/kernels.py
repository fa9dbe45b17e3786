"""Tensor-level front end of the C-ABI kernels (include/mirror_b200.h).

Every function takes torch tensors (device memory owned by PyTorch's caching
allocator), extracts raw pointers / strides and enqueues the kernel on the
current CUDA stream.  There is no CPU or eager fallback: a missing library or a
non-CUDA tensor raises.

``_TEST_BACKEND`` exists for the CPU unit tests only (tests/emu_backend.py
injects a torch re-statement of each entry point so the host logic and the
hand-written backward passes can be checked against the oracle without a GPU).
Product code never sets it.
"""
import ctypes

import torch

from . import _lib

_TEST_BACKEND = None  # set by tests/emu_backend.py only

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
LAUNCHES = [0]  # number of kernels enqueued (bench.py reports it as gpu_launches)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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


def gemm(a, b, *, out_f32=None, out_bf16=None, alpha=1.0, bias=None, act=ACT_NONE, drop_p=0.0, drop_seed=0,
         res=None, gamma=1.0, beta=0.0, split_k=1, simt=False):
    """C[..,M,N] = epilogue(A[..,M,K] @ B[..,N,K]^T); see mirror_gemm_args.

    ``a`` / ``b`` are bf16 views whose last two dims are (M,K) / (N,K); either
    of the two may carry the unit stride, so ``x.transpose(-1,-2)`` views give
    the NN / TN forms without copies.  Up to two leading batch dims.
    """
    if _TEST_BACKEND is not None:
        return _TEST_BACKEND.gemm(a, b, out_f32=out_f32, out_bf16=out_bf16, alpha=alpha, bias=bias, act=act,
                                  drop_p=drop_p, drop_seed=drop_seed, res=res, gamma=gamma, beta=beta, split_k=split_k)
    a4, b4 = _as4(_cuda(a, torch.bfloat16)), _as4(_cuda(b, torch.bfloat16))
    B2, B1, M, K = a4.shape
    N = b4.shape[2]
    if b4.shape != (B2, B1, N, K):
        raise ValueError(f"gemm shape mismatch {tuple(a.shape)} x {tuple(b.shape)}")
    g = _lib.GemmArgs()
    g.a, g.b = a4.data_ptr(), b4.data_ptr()
    g.a_mn_major, g.lda = _major(a4)
    g.b_mn_major, g.ldb = _major(b4)
    g.a_bs1, g.a_bs2, g.b_bs1, g.b_bs2 = a4.stride(1), a4.stride(0), b4.stride(1), b4.stride(0)
    g.M, g.N, g.K, g.batch1, g.batch2 = M, N, K, B1, B2
    g.alpha = alpha
    g.bias = _cuda(bias, torch.float32).data_ptr() if bias is not None else None
    g.act, g.drop_p, g.drop_seed = act, drop_p, drop_seed
    if res is not None:
        r4 = _as4(_cuda(res))
        if r4.shape != (B2, B1, M, N) or r4.stride(-1) != 1:
            raise ValueError("gemm residual must be [..,M,N] with unit column stride")
        g.res, g.res_is_bf16 = r4.data_ptr(), int(r4.dtype == torch.bfloat16)
        if r4.dtype not in (torch.bfloat16, torch.float32):
            raise TypeError("residual must be bf16 or f32")
        g.ldr, g.r_bs1, g.r_bs2 = r4.stride(2), r4.stride(1), r4.stride(0)
    g.gamma, g.beta, g.split_k = gamma, beta, split_k
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
    fn = _lib.lib().mirror_gemm_bf16_simt if simt else _lib.lib().mirror_gemm_bf16
    _lib.check(fn(ctypes.byref(g), _stream()), "gemm")
    LAUNCHES[0] += 1
