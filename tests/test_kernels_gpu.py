"""GPU: attention / LayerNorm kernels through the C ABI against plain torch fp32 references of the same op
(operands rounded to bf16 exactly as the kernel sees them).  Tolerances: bf16 output rounding (2^-8
relative) plus accumulation-order noise."""
import pytest
import torch
import torch.nn.functional as F

from fedcola_b200 import ops

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol, atol, what=""):
    err = (got.float() - ref.float()).abs()
    tol = atol + rtol * ref.float().abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad} mismatches, max err {err.max().item():.4e} (ref max {ref.abs().max().item():.3e})"


def _attn_ref(qkv, B, N, H):
    """Attention.forward between the Linears (mome.py:153-165) on fp32 copies of the bf16 operands."""
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    q = q * 64 ** -0.5
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    attn_b = attn.to(torch.bfloat16).float()          # `.type_as(x)` in the bf16 pipeline
    return (attn_b @ v).transpose(1, 2).reshape(B, N, H * 64), attn


@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (3, 64, 6), (2, 40, 1), (1, 16, 2), (2, 200, 2), (2, 256, 2),
                                   (3, 129, 1), (2, 128, 3), (5, 1, 2), (4, 7, 1)])
def test_attention_forward(B, N, H, cuda):
    torch.manual_seed(0)
    qkv = (torch.randn(B, N, 3, H, 64, device=cuda) * 1.5).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, N, H)
    ref, _ = _attn_ref(qkv, B, N, H)
    # The kernel rounds the un-normalised probabilities exp(s - max) to bf16 and divides the fp32 row of P·V by the
    # fp32 row sum; the reference pipeline in bf16 would round the normalised ones.  Both are a 2^-9 relative
    # rounding of every term of a convex combination of V rows: |error| <= ~2^-8 * max|V| + output rounding.
    vmax = qkv.float().view(B, N, 3, H, 64)[:, :, 2].abs().max().item()
    _close(out, ref, 2 ** -7, 2 ** -8 * vmax, "out")
    x = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    exact = (((x[0] * 0.125) @ x[1].transpose(-2, -1)).softmax(-1) @ x[2]).transpose(1, 2).reshape(B, N, H * 64)
    rel = (out.float() - exact).norm() / exact.norm()
    assert rel < 4e-3, rel                            # aggregate error vs the fp32 op: bf16 output rounding level
    q, k, _ = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref_lse = torch.logsumexp((q * 0.125) @ k.transpose(-2, -1), dim=-1)
    _close(lse, ref_lse, 1e-4, 1e-3, "lse")


@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (3, 64, 6), (2, 40, 1), (2, 200, 2), (2, 128, 2), (1, 256, 1),
                                   (3, 129, 2), (4, 7, 3)])
def test_attention_backward(B, N, H, cuda):
    torch.manual_seed(1)
    qkv = torch.randn(B, N, 3, H, 64, device=cuda).to(torch.bfloat16)
    dout = (torch.randn(B, N, H * 64, device=cuda) * 0.1).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, N, H)
    dbias = torch.ones(3 * H * 64, device=cuda)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, dbias=dbias)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    o = (((q * 0.125) @ k.transpose(-2, -1)).softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * 64)
    o.backward(dout.float())
    # fused qkv bias gradient: dbias += column sums of dQ and dV (the V third is taken exactly, as the column sums of
    # dO through an all-ones row of P^T); the K third is left untouched because it is identically zero (softmax
    # is shift-invariant) — the fp32 autograd value is ~1e-7.  Bar: 2e-2 of the gradient's scale (bf16 mode).
    gb = x.grad.view(B * N, 3, H * 64).sum(0)
    assert gb[1].abs().max().item() < 1e-4 * x.grad.abs().max().item() * (B * N) ** 0.5, "reference K bias gradient is ~0"
    got = dbias.view(3, -1) - 1.0
    for third, name in ((0, "q"), (2, "v")):
        _close(got[third], gb[third], 2e-2, 2e-2 * gb[third].abs().max().item(), f"fused {name} bias gradient")
    assert torch.equal(dbias.view(3, -1)[1], torch.ones(H * 64, device=cuda))
    scale = x.grad.abs().max().item()
    _close(dqkv, x.grad, 3e-2, 1.5e-2 * scale, "dqkv")
    # aggregate error well inside the bf16 budget
    rel = (dqkv.float() - x.grad).norm() / x.grad.norm()
    assert rel < 1e-2, rel


@pytest.mark.parametrize("rows,d,eps", [(1000, 384, 1e-5), (197 * 3, 192, 1e-5), (64, 768, 1e-6), (33, 64, 1e-12)])
def test_layernorm_forward_backward(rows, d, eps, cuda):
    torch.manual_seed(2)
    x = torch.randn(rows, d, device=cuda) * 2 + 0.5
    g = torch.randn(d, device=cuda) * 0.2 + 1
    b = torch.randn(d, device=cuda) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b, eps, bf16_out=False)
    ref = F.layer_norm(x, (d,), g, b, eps)
    _close(y, ref, 1e-5, 1e-5, "ln fwd fp32")
    yb, _, _ = ops.layernorm_fwd(x, g, b, eps, bf16_out=True)
    _close(yb, ref, 2 ** -8, 1e-3, "ln fwd bf16")
    # backward with accumulate + scaled bf16 copy
    dy = torch.randn(rows, d, device=cuda)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (d,), gr, br, eps).backward(dy)
    dx0 = torch.randn(rows, d, device=cuda)
    dx = dx0.clone()
    group = 7
    scale = torch.rand((rows + group - 1) // group, device=cuda) + 0.5
    dxs = torch.empty(rows, d, dtype=torch.bfloat16, device=cuda)
    dgamma, dbeta = torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
    colsum = torch.full((d,), 2.0, device=cuda)
    ops.layernorm_bwd(dy, x, mean, rstd, g, dx, True, dxs=dxs, row_scale=scale, rows_per_group=group, dgamma=dgamma,
                      dbeta=dbeta, dxs_colsum=colsum)
    _close(colsum, 2.0 + dxs.float().sum(0), 1e-4, 1e-3 * rows ** 0.5, "fused bias gradient (colsum of dxs)")
    _close(dx, dx0 + xr.grad, 1e-4, 1e-4, "ln bwd dx")
    _close(dxs, (dx0 + xr.grad) * scale.repeat_interleave(group)[:rows, None], 2 ** -8, 1e-3, "ln bwd dxs")
    _close(dgamma, gr.grad, 1e-3, 1e-3 * rows ** 0.5, "dgamma")
    _close(dbeta, br.grad, 1e-3, 1e-3 * rows ** 0.5, "dbeta")
    # bf16 dy path
    dx2 = torch.zeros(rows, d, device=cuda)
    ops.layernorm_bwd(dy.to(torch.bfloat16), x, mean, rstd, g, dx2, False)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (d,), g, b, eps).backward(dy.to(torch.bfloat16).float())
    _close(dx2, xr2.grad, 1e-4, 1e-4, "ln bwd bf16 dy")


@pytest.mark.parametrize("rows,d,accumulate,bf16_dy", [(22064, 384, True, True), (7168, 384, True, True), (6000, 768, True, True),
                                                       (5003, 384, False, False), (18912, 768, True, True)])
def test_layernorm_backward_streamed_large(rows, d, accumulate, bf16_dy, cuda):
    """Launches big enough for the bulk-copy ring version of the LayerNorm backward (>= 4 736 rows): the bench's own
    shapes (112 x 197 / 112 x 64 rows at d=384, 96 x 197 at d=768), ragged row counts, with and without the running-dx
    accumulation, against torch's autograd."""
    torch.manual_seed(3)
    x = torch.randn(rows, d, device=cuda) * 2 + 0.5
    g = torch.randn(d, device=cuda) * 0.2 + 1
    b = torch.randn(d, device=cuda) * 0.1
    _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, bf16_out=True)
    dy = torch.randn(rows, d, device=cuda)
    if bf16_dy:
        dy = dy.to(torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy.float())
    dx0 = torch.randn(rows, d, device=cuda)
    dx = dx0.clone()
    group = 197
    scale = torch.rand((rows + group - 1) // group, device=cuda) + 0.5
    dxs = torch.empty(rows, d, dtype=torch.bfloat16, device=cuda)
    dgamma, dbeta = torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
    colsum = torch.zeros(d, device=cuda)
    ops.layernorm_bwd(dy, x, mean, rstd, g, dx, accumulate, dxs=dxs, row_scale=scale, rows_per_group=group, dgamma=dgamma,
                      dbeta=dbeta, dxs_colsum=colsum)
    want = (dx0 + xr.grad) if accumulate else xr.grad
    _close(dx, want, 1e-4, 1e-4, "dx")
    _close(dxs, want * scale.repeat_interleave(group)[:rows, None], 2 ** -8, 1e-3, "dxs")
    _close(colsum, dxs.float().sum(0), 1e-4, 2e-3 * rows ** 0.5, "colsum of dxs")
    _close(dgamma, gr.grad, 1e-3, 2e-3 * rows ** 0.5, "dgamma")
    _close(dbeta, br.grad, 1e-3, 2e-3 * rows ** 0.5, "dbeta")
