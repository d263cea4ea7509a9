"""Thin Python wrappers (ctypes) over the per-kernel C-ABI entry points — used by the parity tests and by
code paths that need a single op.  The training hot loop does NOT go through here: it makes one native
call per step (runtime.py -> csrc/mat_driver.cu)."""
import contextlib

import torch

from . import _lib
from ._lib import c_f, c_int, c_ll, c_vp, ptr

EPI_BF16, EPI_GELU, EPI_RESID, EPI_MULAUX, EPI_F32, EPI_ATOMIC_F32, EPI_PATCH = range(7)


@contextlib.contextmanager
def grid_cap(max_ctas):
    """Cap the CTA count of the persistent kernels inside the block (fc_set_grid_cap): small problems then run
    many tiles / items per CTA — the regime the bench shapes run in.  Process-wide (tests are single-threaded)."""
    L = _lib.lib()
    prev = L.fc_set_grid_cap(c_int(int(max_ctas)))
    try:
        yield
    finally:
        L.fc_set_grid_cap(c_int(prev))


def _dev(t):
    _lib.require_cuda(t)
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def gemm_bf16(A, B, epi, out, *, a_mn=False, b_mn=False, out2=None, bias=None, resid=None, row_scale=None,
              rows_per_group=0, aux=None, pos=None, patches=0, alpha=1.0, splits=1, colsum=None):
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
                                 c_int(patches), c_f(alpha), c_int(splits), ptr(colsum), c_int(dev), _lib.stream_ptr(A.device))
    _lib.check(rc, "fc_gemm_bf16")
    return out


def _ptr_table(ts):
    """host array of device pointers (None -> NULL table)"""
    import ctypes
    if ts is None:
        return None
    return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


def gemm_bf16_grouped(As, Bs, epi, outs, *, a_mn=False, b_mn=False, out2s=None, biases=None, resids=None,
                      row_scales=None, rows_per_group=0, auxs=None, poss=None, patches=0, alpha=1.0, splits=1, colsums=None):
    """The same GEMM for len(As) independent operand sets (same shapes / strides) in ONE launch (fc_gemm_bf16_grouped)."""
    G = len(As)
    A, B, out = As[0], Bs[0], outs[0]
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N = B.shape[1] if b_mn else B.shape[0]
    for a, b, o in zip(As, Bs, outs):
        assert a.shape == A.shape and b.shape == B.shape and o.shape == out.shape
        assert a.stride() == A.stride() and b.stride() == B.stride() and o.stride() == out.stride()
    dev = _dev(A)
    ldo = out.stride(0) if epi != EPI_PATCH else out.shape[-1]
    rc = _lib.lib().fc_gemm_bf16_grouped(c_int(G), c_int(M), c_int(N), c_int(K), _ptr_table(As), c_ll(A.stride(0)),
                                         c_int(int(a_mn)), _ptr_table(Bs), c_ll(B.stride(0)), c_int(int(b_mn)), c_int(epi),
                                         _ptr_table(outs), _ptr_table(out2s), c_ll(ldo), _ptr_table(biases),
                                         _ptr_table(resids), _ptr_table(row_scales), c_int(rows_per_group), _ptr_table(auxs),
                                         _ptr_table(poss), c_int(patches), c_f(alpha), c_int(splits), _ptr_table(colsums),
                                         c_int(dev), _lib.stream_ptr(A.device))
    _lib.check(rc, "fc_gemm_bf16_grouped")
    return outs


def _st(t):
    return _lib.stream_ptr(t.device)


def attention_fwd(qkv, B, N, H, want_lse=True):
    """qkv bf16 [B,N,3,H,64] -> (out bf16 [B,N,H*64], lse fp32 [B,H,N])"""
    out = torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device) if want_lse else None
    rc = _lib.lib().fc_attention_fwd(ptr(qkv), ptr(out), ptr(lse), c_int(B), c_int(N), c_int(H), c_int(64),
                                     c_int(_dev(qkv)), _st(qkv))
    _lib.check(rc, "fc_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, B, N, H, dbias=None):
    dqkv = torch.empty_like(qkv)
    rc = _lib.lib().fc_attention_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), ptr(dbias), c_int(B), c_int(N), c_int(H),
                                     c_int(64), c_int(_dev(qkv)), _st(qkv))
    _lib.check(rc, "fc_attention_bwd")
    return dqkv


def attention_fwd_grouped(qkvs, B, N, H):
    """The same attention for several clients' qkv tensors in ONE launch (fc_attention_fwd_grouped)."""
    outs = [torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=q.device) for q in qkvs]
    lses = [torch.empty(B, H, N, dtype=torch.float32, device=q.device) for q in qkvs]
    rc = _lib.lib().fc_attention_fwd_grouped(c_int(len(qkvs)), _ptr_table(qkvs), _ptr_table(outs), _ptr_table(lses), c_int(B),
                                             c_int(N), c_int(H), c_int(64), c_int(_dev(qkvs[0])), _st(qkvs[0]))
    _lib.check(rc, "fc_attention_fwd_grouped")
    return outs, lses


def attention_bwd_grouped(qkvs, outs, douts, lses, B, N, H, dbiases=None):
    dqkvs = [torch.empty_like(q) for q in qkvs]
    rc = _lib.lib().fc_attention_bwd_grouped(c_int(len(qkvs)), _ptr_table(qkvs), _ptr_table(outs), _ptr_table(douts),
                                             _ptr_table(lses), _ptr_table(dqkvs), _ptr_table(dbiases), c_int(B), c_int(N),
                                             c_int(H), c_int(64), c_int(_dev(qkvs[0])), _st(qkvs[0]))
    _lib.check(rc, "fc_attention_bwd_grouped")
    return dqkvs


def layernorm_fwd(x, gamma, beta, eps, bf16_out=True):
    rows, d = x.shape
    y = torch.empty(rows, d, dtype=torch.bfloat16 if bf16_out else torch.float32, device=x.device)
    mean = torch.empty(rows, device=x.device)
    rstd = torch.empty(rows, device=x.device)
    rc = _lib.lib().fc_layernorm_fwd(ptr(x), c_ll(x.stride(0)), ptr(gamma), ptr(beta), c_f(eps),
                                     ptr(y if bf16_out else None), ptr(None if bf16_out else y), ptr(mean), ptr(rstd),
                                     c_int(rows), c_int(d), c_int(_dev(x)), _st(x))
    _lib.check(rc, "fc_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, accumulate, dxs=None, row_scale=None, rows_per_group=0,
                  dgamma=None, dbeta=None, dxs_colsum=None):
    rows, d = x.shape
    rc = _lib.lib().fc_layernorm_bwd(ptr(dy), c_int(int(dy.dtype == torch.bfloat16)), c_ll(dy.stride(0)), ptr(x),
                                     c_ll(x.stride(0)), ptr(mean), ptr(rstd), ptr(gamma), ptr(dx), c_ll(dx.stride(0)),
                                     c_int(int(accumulate)), ptr(dxs), c_ll(dxs.stride(0) if dxs is not None else d),
                                     ptr(row_scale), c_int(rows_per_group), ptr(dgamma), ptr(dbeta), ptr(dxs_colsum), c_int(rows),
                                     c_int(d), c_int(_dev(x)), _st(x))
    _lib.check(rc, "fc_layernorm_bwd")
    return dx


def layernorm_bwd_grouped(dys, xs, means, rstds, gammas, dxs_out, accumulate, dxs=None, row_scales=None, rows_per_group=0,
                          dgammas=None, dbetas=None, dxs_colsums=None):
    """The same LayerNorm backward for len(xs) operand sets of one shape in ONE launch (fc_layernorm_bwd_grouped)."""
    rows, d = xs[0].shape
    dy, x, dx = dys[0], xs[0], dxs_out[0]
    rc = _lib.lib().fc_layernorm_bwd_grouped(
        c_int(len(xs)), _ptr_table(dys), c_int(int(dy.dtype == torch.bfloat16)), c_ll(dy.stride(0)), _ptr_table(xs),
        c_ll(x.stride(0)), _ptr_table(means), _ptr_table(rstds), _ptr_table(gammas), _ptr_table(dxs_out), c_ll(dx.stride(0)),
        c_int(int(accumulate)), _ptr_table(dxs), c_ll(dxs[0].stride(0) if dxs is not None else d), _ptr_table(row_scales),
        c_int(rows_per_group), _ptr_table(dgammas), _ptr_table(dbetas), _ptr_table(dxs_colsums), c_int(rows), c_int(d),
        c_int(_dev(x)), _st(x))
    _lib.check(rc, "fc_layernorm_bwd_grouped")
    return dxs_out


# ---- fp32-accurate validation mode (csrc/precise.cu, fc_gemm_split) ------------------------------------
def split_bf16(x, row_scale=None, rows_per_group=0):
    """fp32 tensor -> (hi, lo) bf16 pair with x ~= hi + lo (16 mantissa bits)."""
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    rc = _lib.lib().fc_split_bf16(ptr(x), ptr(hi), ptr(lo), c_ll(x.numel()), c_int(x.shape[-1]), ptr(row_scale),
                                  c_int(rows_per_group), c_int(_dev(x)), _st(x))
    _lib.check(rc, "fc_split_bf16")
    return hi, lo


def gemm_split(A, B, epi, out, *, a_mn=False, b_mn=False, bias=None, resid=None, row_scale=None, rows_per_group=0,
               alpha=1.0, splits=1):
    """out (+)= A · Bᵀ for fp32 A, B through split bf16 operands (3 tcgen05 passes): fp32-accurate."""
    assert A.dtype == torch.float32 and B.dtype == torch.float32
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N = B.shape[1] if b_mn else B.shape[0]
    ah, al = split_bf16(A)
    bh, bl = split_bf16(B)
    rc = _lib.lib().fc_gemm_split(c_int(M), c_int(N), c_int(K), ptr(ah), ptr(al), c_ll(A.stride(0)), c_int(int(a_mn)), ptr(bh),
                                  ptr(bl), c_ll(B.stride(0)), c_int(int(b_mn)), c_int(epi), ptr(out), c_ll(out.stride(0)),
                                  ptr(bias), ptr(resid), ptr(row_scale), c_int(rows_per_group), ptr(None), c_int(0),
                                  c_f(alpha), c_int(splits), c_int(_dev(A)), _st(A))
    _lib.check(rc, "fc_gemm_split")
    return out


def attention_f32(qkv, B, N, H, dout=None):
    """fp32 FMA attention of the validation mode: returns (out, lse[, dqkv])."""
    out = torch.empty(B, N, H * 64, dtype=torch.float32, device=qkv.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device)
    L = _lib.lib()
    _lib.check(L.fc_attention_f32_fwd(ptr(qkv), ptr(out), ptr(lse), c_int(B), c_int(N), c_int(H), c_int(_dev(qkv)), _st(qkv)),
               "fc_attention_f32_fwd")
    if dout is None:
        return out, lse
    dqkv = torch.zeros_like(qkv)
    _lib.check(L.fc_attention_f32_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), c_int(B), c_int(N), c_int(H),
                                      c_int(_dev(qkv)), _st(qkv)), "fc_attention_f32_bwd")
    return out, lse, dqkv
