"""Thin Python wrappers (ctypes) over the per-kernel C-ABI entry points — used by the parity tests and by
code paths that need a single op.  The training hot loop does NOT go through here: it makes one native
call per step (runtime.py -> csrc/mat_driver.cu)."""
import torch

from . import _lib
from ._lib import c_f, c_int, c_ll, c_vp, ptr

EPI_BF16, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_F32, EPI_ATOMIC_F32, EPI_PATCH = range(7)


def _dev(t):
    _lib.require_cuda(t)
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def gemm_bf16(A, B, epi, out, *, a_mn=False, b_mn=False, out2=None, bias=None, resid=None, row_scale=None,
              rows_per_group=0, aux=None, pos=None, patches=0, alpha=1.0, splits=1):
    """out[M,N] (+)= A · Bᵀ with the fused epilogue `epi` (see include/fedcola_b200.h).
    A: [M,K] (or [K,M] if a_mn), B: [N,K] (or [K,N] if b_mn); bf16, row-major, last dim contiguous."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N, K2 = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    assert K == K2, (A.shape, B.shape)
    dev = _dev(A)
    ldo = out.stride(0) if epi != EPI_PATCH else out.shape[-1]
    rc = _lib.lib().fc_gemm_bf16(c_int(M), c_int(N), c_int(K), ptr(A), c_ll(A.stride(0)), c_int(int(a_mn)), ptr(B),
                                 c_ll(B.stride(0)), c_int(int(b_mn)), c_int(epi), ptr(out), ptr(out2), c_ll(ldo),
                                 ptr(bias), ptr(resid), ptr(row_scale), c_int(rows_per_group), ptr(aux), ptr(pos),
                                 c_int(patches), c_f(alpha), c_int(splits), c_int(dev), _lib.stream_ptr(A.device))
    _lib.check(rc, "fc_gemm_bf16")
    return out
