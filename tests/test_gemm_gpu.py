"""GPU: the tcgen05 GEMM (csrc/gemm.cu) through the C ABI against a plain torch fp32 reference of the
same op on the same bf16-rounded operands.  Tolerance: fp32 accumulation order only (1e-3 relative of the
row scale) for fp32 outputs, one bf16 ulp (2^-8 relative) for bf16 outputs."""
import pytest
import torch
import torch.nn.functional as F

from fedcola_b200 import ops

pytestmark = pytest.mark.gpu


def _mk(rows, cols, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(dev).to(torch.bfloat16)


def _close(got, ref, rtol, atol):
    err = (got.float() - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{bad} mismatches, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


SHAPES = [(256, 128, 64), (200, 192, 192), (1000, 384, 384), (197 * 5, 1152, 384), (333, 1536, 384), (130, 64, 1536)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_nt_plain_f32(M, N, K, cuda):
    A, B = _mk(M, K, cuda, 1), _mk(N, K, cuda, 2)
    out = torch.full((M, N), float("nan"), device=cuda)
    ops.gemm_bf16(A, B, ops.EPI_F32, out)
    _close(out, A.float() @ B.float().t(), 1e-3, 1e-2)


@pytest.mark.parametrize("M,N,K", SHAPES[:4])
def test_nt_bias_bf16(M, N, K, cuda):
    A, B = _mk(M, K, cuda, 3), _mk(N, K, cuda, 4)
    bias = torch.randn(N, device=cuda)
    out = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm_bf16(A, B, ops.EPI_BF16, out, bias=bias)
    _close(out, A.float() @ B.float().t() + bias, 2 ** -7, 2e-2)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (384, 1152, 1000), (192, 192, 333 * 8), (64, 256, 197 * 8)])
def test_operand_majors(a_mn, b_mn, M, N, K, cuda):
    A = _mk(K, M, cuda, 5) if a_mn else _mk(M, K, cuda, 5)
    B = _mk(K, N, cuda, 6) if b_mn else _mk(N, K, cuda, 6)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    out = torch.full((M, N), float("nan"), device=cuda)
    ops.gemm_bf16(A, B, ops.EPI_F32, out, a_mn=a_mn, b_mn=b_mn)
    _close(out, Af @ Bf.t(), 1e-3, 3e-2)


@pytest.mark.parametrize("splits", [1, 3, 7])
def test_tn_split_k_atomic(splits, cuda):
    """dW = dY^T X accumulated into an existing gradient (both operands MN-major)."""
    T, No, Ko = 197 * 6, 384, 192
    dY, X = _mk(T, No, cuda, 7, 0.1), _mk(T, Ko, cuda, 8)
    out = torch.ones(No, Ko, device=cuda)
    ops.gemm_bf16(dY, X, ops.EPI_ATOMIC_F32, out, a_mn=True, b_mn=True, splits=splits, alpha=0.5)
    _close(out, 1.0 + 0.5 * (dY.float().t() @ X.float()), 1e-3, 1e-2)


def test_gelu_epilogue(cuda):
    """fc1 forward: out = gelu'(x) (saved for the backward), out2 = gelu(x), x = A W^T + b (exact-erf GELU)."""
    M, N, K = 500, 1536, 384
    A, B = _mk(M, K, cuda, 9, 0.5), _mk(N, K, cuda, 10, 0.1)
    bias = torch.randn(N, device=cuda) * 0.1
    dact = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
    act = torch.zeros_like(dact)
    ops.gemm_bf16(A, B, ops.EPI_GELU, dact, out2=act, bias=bias)
    ref = (A.float() @ B.float().t() + bias).requires_grad_(True)
    y = F.gelu(ref)
    y.sum().backward()
    _close(act, y.detach(), 2 ** -7, 1e-2)
    _close(dact, ref.grad, 2 ** -7, 1e-2)


def test_gelu_matches_exact_erf_over_range(cuda):
    """The rational erf of the epilogue against torch's exact-erf GELU and its derivative on [-12, 12]:
    acc[m, n] = A[m, 0] * B[n, 0] = x_n for every row."""
    M, N, K = 128, 1024, 64
    A = torch.zeros(M, K, device=cuda, dtype=torch.bfloat16)
    B = torch.zeros(N, K, device=cuda, dtype=torch.bfloat16)
    A[:, 0] = 1.0
    B[:, 0] = torch.linspace(-12, 12, N, device=cuda).to(torch.bfloat16)
    dact = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
    act = torch.zeros_like(dact)
    ops.gemm_bf16(A, B, ops.EPI_GELU, dact, out2=act)
    x = B[:, 0].float().requires_grad_(True)
    y = F.gelu(x)
    y.sum().backward()
    for r in (0, M - 1):
        _close(act[r], y.detach(), 2 ** -8, 1e-6)
        _close(dact[r], x.grad, 2 ** -8, 1e-6)


def test_mulaux_epilogue_with_bias_grad(cuda):
    """fc2 backward: d_h = (dY W2) * gelu'(pre) with the fc1 bias gradient (column sums of d_h) fused."""
    M, N, K = 300, 1536, 384
    A, B = _mk(M, K, cuda, 11, 0.5), _mk(K, N, cuda, 12, 0.1)       # B MN-major: W2 is stored [d, 4d]
    aux = _mk(M, N, cuda, 13)
    out = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
    colsum = torch.ones(N, device=cuda)
    ops.gemm_bf16(A, B, ops.EPI_MULAUX, out, b_mn=True, aux=aux, colsum=colsum)
    ref = (A.float() @ B.float()) * aux.float()
    _close(out, ref, 2 ** -7, 1e-2)
    _close(colsum, 1.0 + out.float().sum(0), 1e-4, 1e-3)          # sums the bf16 values actually stored
    # ragged N (not a multiple of the 32-column epilogue chunk)
    Bt = _mk(K, 200, cuda, 14, 0.1)
    aux2 = _mk(M, 200, cuda, 15)
    out2 = torch.zeros(M, 200, device=cuda, dtype=torch.bfloat16)
    cs2 = torch.zeros(200, device=cuda)
    ops.gemm_bf16(A, Bt, ops.EPI_MULAUX, out2, b_mn=True, aux=aux2, colsum=cs2)
    _close(out2, (A.float() @ Bt.float()) * aux2.float(), 2 ** -7, 1e-2)
    _close(cs2, out2.float().sum(0), 1e-4, 1e-3)


def test_residual_droppath_epilogue(cuda):
    Bsz, Ntok, d, K = 6, 197, 384, 1536
    M = Bsz * Ntok
    A, W = _mk(M, K, cuda, 14, 0.5), _mk(d, K, cuda, 15, 0.05)
    bias = torch.randn(d, device=cuda) * 0.1
    x = torch.randn(M, d, device=cuda)
    keep = torch.tensor([0.0, 1 / 0.9, 1 / 0.9, 0.0, 1 / 0.9, 1 / 0.9], device=cuda)
    out = torch.empty_like(x)
    ops.gemm_bf16(A, W, ops.EPI_RESID, out, bias=bias, resid=x, row_scale=keep, rows_per_group=Ntok)
    ref = x + keep.repeat_interleave(Ntok)[:, None] * (A.float() @ W.float().t() + bias)
    _close(out, ref, 1e-3, 1e-2)
    # in place on the residual stream
    x2 = x.clone()
    ops.gemm_bf16(A, W, ops.EPI_RESID, x2, bias=bias, resid=x2)
    _close(x2, x + A.float() @ W.float().t() + bias, 1e-3, 1e-2)


def test_patch_epilogue(cuda):
    Bsz, P, d, K = 3, 196, 192, 768
    A, W = _mk(Bsz * P, K, cuda, 16), _mk(d, K, cuda, 17, 0.05)
    bias = torch.randn(d, device=cuda) * 0.1
    pos = torch.randn(P + 1, d, device=cuda)
    out = torch.zeros(Bsz, P + 1, d, device=cuda)
    ops.gemm_bf16(A, W, ops.EPI_PATCH, out, bias=bias, pos=pos, patches=P)
    ref = (A.float() @ W.float().t() + bias).view(Bsz, P, d) + pos[1:]
    _close(out[:, 1:], ref, 1e-3, 1e-2)
    assert torch.count_nonzero(out[:, 0]) == 0


def test_unbuilt_combination_is_refused(cuda):
    """Only the (operand majors x epilogue) pairs the round uses are instantiated; others fail loudly."""
    A, B = _mk(128, 64, cuda, 20), _mk(128, 64, cuda, 21)
    out = torch.zeros(128, 128, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="not built"):
        ops.gemm_bf16(A, B, ops.EPI_MULAUX, out, aux=out)
