"""GPU: the persistent tcgen05 kernels in the regime the bench runs them in — MANY tiles / items per CTA.

`grid = min(#SMs, tiles)`, so small test problems normally give every CTA at most one tile and the per-CTA
state machines (TMEM double-buffer parity `(it >> 1) & 1`, the TMA ring phase carried across tile boundaries, the
attention item ring with its deferred accumulator flush) never advance.  Two ways to exercise them under a checker:
  * `ops.grid_cap(n)` (fc_set_grid_cap) forces a small grid on a small problem: >= 3 tiles per CTA for every
    (operand majors x epilogue) combination that is instantiated, and >= 3 items per CTA for attention;
  * the real shapes of the bench (ViT-S B=112: T = 22 064 / 7 168 tokens; ViT-B B=96) with the natural grid:
    ~1 000 tiles over 148 CTAs, 672 / 1 152 attention items.
Checked against a plain torch fp32 reference of the same op on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from fedcola_b200 import ops

pytestmark = pytest.mark.gpu


def _mk(rows, cols, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(dev).to(torch.bfloat16)


def _close(got, ref, rtol, atol, what=""):
    err = (got.float() - ref.float()).abs()
    tol = atol + rtol * ref.float().abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad} mismatches, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


def _run_epilogue(epi, a_mn, b_mn, M, N, K, dev, seed=0, splits=1):
    """One launch of (majors, epilogue) and its fp32 torch reference. Returns [(got, ref, rtol, atol, name)]."""
    A = _mk(K, M, dev, seed + 1, 0.5) if a_mn else _mk(M, K, dev, seed + 1, 0.5)
    B = _mk(K, N, dev, seed + 2, 0.1) if b_mn else _mk(N, K, dev, seed + 2, 0.1)
    acc = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    bias = torch.randn(N, device=dev) * 0.1
    if epi == ops.EPI_F32:
        out = torch.full((M, N), float("nan"), device=dev)
        ops.gemm_bf16(A, B, epi, out, a_mn=a_mn, b_mn=b_mn, bias=bias)
        return [(out, acc + bias, 1e-3, 1e-2, "f32")]
    if epi == ops.EPI_BF16:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm_bf16(A, B, epi, out, a_mn=a_mn, b_mn=b_mn, bias=bias)
        return [(out, acc + bias, 2 ** -7, 2e-2, "bf16")]
    if epi == ops.EPI_GELU:
        dact = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        act = torch.zeros_like(dact)
        ops.gemm_bf16(A, B, epi, dact, out2=act, bias=bias)
        x = (acc + bias).requires_grad_(True)
        y = F.gelu(x)
        y.sum().backward()
        return [(act, y.detach(), 2 ** -7, 1e-2, "gelu"), (dact, x.grad, 2 ** -7, 1e-2, "gelu'")]
    if epi == ops.EPI_RESID:
        rows_per_group = 50
        x = torch.randn(M, N, device=dev)
        keep = (torch.rand((M + rows_per_group - 1) // rows_per_group, device=dev) > 0.3).float() / 0.7
        out = torch.empty_like(x)
        ops.gemm_bf16(A, B, epi, out, bias=bias, resid=x, row_scale=keep, rows_per_group=rows_per_group)
        ref = x + keep.repeat_interleave(rows_per_group)[:M, None] * (acc + bias)
        return [(out, ref, 1e-3, 1e-2, "resid")]
    if epi == ops.EPI_MULAUX:
        aux = _mk(M, N, dev, seed + 3)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        colsum = torch.ones(N, device=dev)
        ops.gemm_bf16(A, B, epi, out, a_mn=a_mn, b_mn=b_mn, aux=aux, colsum=colsum)
        return [(out, acc * aux.float(), 2 ** -7, 1e-2, "mulaux"),
                (colsum, 1.0 + out.float().sum(0), 1e-4, 1e-3 * M ** 0.5, "mulaux colsum")]
    if epi == ops.EPI_ATOMIC_F32:
        out = torch.ones(M, N, device=dev)
        ops.gemm_bf16(A, B, epi, out, a_mn=a_mn, b_mn=b_mn, splits=splits, alpha=0.5)
        return [(out, 1.0 + 0.5 * acc, 1e-3, 1e-2 * max(1.0, (K / 4096) ** 0.5), "atomic")]
    raise AssertionError(epi)


# every (a_mn, b_mn, epilogue) instantiation of csrc/gemm.cu::launch_bn (PATCH has its own test below)
COMBOS = [(0, 0, ops.EPI_BF16), (0, 0, ops.EPI_GELU), (0, 0, ops.EPI_RESID), (0, 0, ops.EPI_F32),
          (0, 1, ops.EPI_BF16), (0, 1, ops.EPI_MULAUX), (0, 1, ops.EPI_F32),
          (1, 1, ops.EPI_ATOMIC_F32), (1, 1, ops.EPI_F32), (1, 0, ops.EPI_F32)]


@pytest.mark.parametrize("cap", [3, 7])
@pytest.mark.parametrize("a_mn,b_mn,epi", COMBOS)
@pytest.mark.parametrize("M,N,K", [(1000, 768, 384), (904, 1152, 192), (1304, 384, 1536)])
def test_gemm_many_tiles_per_cta(M, N, K, a_mn, b_mn, epi, cap, cuda):
    """8-11 row tiles x 2-9 column tiles on 3 / 7 CTAs: 5-30 tiles per CTA, ragged M edge, every tile width."""
    with ops.grid_cap(cap):
        results = _run_epilogue(epi, bool(a_mn), bool(b_mn), M, N, K, cuda, seed=10 * epi + cap,
                                splits=3 if epi == ops.EPI_ATOMIC_F32 else 1)
    for got, ref, rtol, atol, name in results:
        _close(got, ref, rtol, atol, name)


@pytest.mark.parametrize("cap", [2, 5])
def test_patch_epilogue_many_tiles_per_cta(cap, cuda):
    Bsz, P, d, K = 7, 196, 384, 768
    A, W = _mk(Bsz * P, K, cuda, 16), _mk(d, K, cuda, 17, 0.05)
    bias = torch.randn(d, device=cuda) * 0.1
    pos = torch.randn(P + 1, d, device=cuda)
    out = torch.zeros(Bsz, P + 1, d, device=cuda)
    with ops.grid_cap(cap):
        ops.gemm_bf16(A, W, ops.EPI_PATCH, out, bias=bias, pos=pos, patches=P)
    ref = (A.float() @ W.float().t() + bias).view(Bsz, P, d) + pos[1:]
    _close(out[:, 1:], ref, 1e-3, 1e-2, "patch")
    assert torch.count_nonzero(out[:, 0]) == 0


def _vit_gemms(T, d):
    """The GEMMs of one transformer block, forward and backward, as (name, epi, a_mn, b_mn, M, N, K, splits)."""
    h = 4 * d
    return [("fwd qkv", ops.EPI_BF16, 0, 0, T, 3 * d, d, 1), ("fwd proj", ops.EPI_RESID, 0, 0, T, d, d, 1),
            ("fwd fc1", ops.EPI_GELU, 0, 0, T, h, d, 1), ("fwd fc2", ops.EPI_RESID, 0, 0, T, d, h, 1),
            ("dX fc2", ops.EPI_MULAUX, 0, 1, T, h, d, 1), ("dX fc1", ops.EPI_BF16, 0, 1, T, d, h, 1),
            ("dX qkv", ops.EPI_BF16, 0, 1, T, d, 3 * d, 1), ("dX proj", ops.EPI_BF16, 0, 1, T, d, d, 1),
            ("dW qkv", ops.EPI_ATOMIC_F32, 1, 1, 3 * d, d, T, 0), ("dW fc1", ops.EPI_ATOMIC_F32, 1, 1, h, d, T, 0),
            ("dW fc2", ops.EPI_ATOMIC_F32, 1, 1, d, h, T, 0), ("dW proj", ops.EPI_ATOMIC_F32, 1, 1, d, d, T, 0)]


REAL = [(f"{tag} {g[0]}", g) for tag, T, d in (("ViT-S img B=112", 112 * 197, 384), ("ViT-S txt B=112", 112 * 64, 384),
                                              ("ViT-B img B=96", 96 * 197, 768), ("ViT-B txt B=96", 96 * 64, 768))
        for g in _vit_gemms(T, d)]


@pytest.mark.parametrize("name,g", REAL, ids=[r[0].replace(" ", "_") for r in REAL])
def test_gemm_bench_shapes(name, g, cuda):
    """The exact launches of the bench round (BASELINE configs[1] ViT-S, configs[3] ViT-B): ~200-1 600 tiles over 148 SMs."""
    _, epi, a_mn, b_mn, M, N, K, splits = g
    for got, ref, rtol, atol, what in _run_epilogue(epi, bool(a_mn), bool(b_mn), M, N, K, cuda, seed=3, splits=splits):
        _close(got, ref, rtol, atol, f"{name} {what}")


# ---- attention -------------------------------------------------------------------------------------------
def _attn_check(B, N, H, dev, seed):
    torch.manual_seed(seed)
    qkv = torch.randn(B, N, 3, H, 64, device=dev).to(torch.bfloat16)
    dout = (torch.randn(B, N, H * 64, device=dev) * 0.1).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, N, H)
    dbias = torch.zeros(3 * H * 64, device=dev)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, dbias=dbias)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q * 0.125) @ k.transpose(-2, -1)
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * 64)
    o.backward(dout.float())
    rel = (out.float() - o.detach()).norm() / o.detach().norm()
    assert rel < 4e-3, ("attention forward", rel.item())
    vmax = qkv.float()[:, :, 2].abs().max().item()
    _close(out, o.detach(), 2 ** -7, 2 ** -7 * vmax, "attention out")
    _close(lse, torch.logsumexp(s.detach(), dim=-1), 1e-4, 1e-3, "lse")
    scale = x.grad.abs().max().item()
    _close(dqkv, x.grad, 3e-2, 1.5e-2 * scale, "dqkv")
    relb = (dqkv.float() - x.grad).norm() / x.grad.norm()
    assert relb < 1e-2, ("attention backward", relb.item())
    gb = x.grad.view(B * N, 3, H * 64).sum(0)
    for third, nm in ((0, "q"), (2, "v")):
        _close(dbias.view(3, -1)[third], gb[third], 2e-2, 2e-2 * gb[third].abs().max().item(), f"{nm} bias gradient")
    assert torch.count_nonzero(dbias.view(3, -1)[1]) == 0


@pytest.mark.parametrize("cap", [2, 5])
@pytest.mark.parametrize("B,N,H", [(4, 197, 3), (6, 64, 3), (5, 40, 2), (3, 256, 2), (4, 129, 2), (7, 16, 3)])
def test_attention_many_items_per_cta(B, N, H, cap, cuda):
    """10-18 (sample, head) items on 2 / 5 CTAs: the operand ring, the S/O double buffers and the deferred
    dV/dK/dQ flush all wrap several times."""
    with ops.grid_cap(cap):
        _attn_check(B, N, H, cuda, seed=cap)


@pytest.mark.parametrize("B,N,H", [(112, 197, 6), (112, 64, 6), (96, 197, 12), (96, 64, 12), (112, 40, 6)])
def test_attention_bench_shapes(B, N, H, cuda):
    """The bench's own launches: 672 (ViT-S) / 1 152 (ViT-B) items over 148 SMs."""
    _attn_check(B, N, H, cuda, seed=9)


# ---- lockstep client groups: one launch for several clients' operands ------------------------------------
@pytest.mark.parametrize("cap", [0, 5])
@pytest.mark.parametrize("a_mn,b_mn,epi", [(0, 0, ops.EPI_BF16), (0, 0, ops.EPI_GELU), (0, 0, ops.EPI_RESID),
                                           (0, 1, ops.EPI_MULAUX), (1, 1, ops.EPI_ATOMIC_F32)])
def test_grouped_gemm_equals_single_launches(a_mn, b_mn, epi, cap, cuda):
    """fc_gemm_bf16_grouped over 3 operand sets == 3 single launches, bit for bit (the split-K accumulation: to fp32
    reordering), including with many tiles per CTA."""
    G, M, N, K = 3, 1000, 768, 384
    As = [_mk(K, M, cuda, 100 + g, 0.5) if a_mn else _mk(M, K, cuda, 100 + g, 0.5) for g in range(G)]
    Bs = [_mk(K, N, cuda, 200 + g, 0.1) if b_mn else _mk(N, K, cuda, 200 + g, 0.1) for g in range(G)]
    biases = [torch.randn(N, device=cuda) * 0.1 for _ in range(G)]
    kw1, kwg = {}, {}
    if epi in (ops.EPI_BF16, ops.EPI_GELU, ops.EPI_MULAUX):
        mk_out = lambda: torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)     # noqa: E731
    else:
        mk_out = lambda: torch.ones(M, N, device=cuda)                            # noqa: E731
    single, grouped = [mk_out() for _ in range(G)], [mk_out() for _ in range(G)]
    extra_s, extra_g = [None] * G, None
    if epi == ops.EPI_GELU:
        extra_s, extra_g = [mk_out() for _ in range(G)], [mk_out() for _ in range(G)]
    resid = [torch.randn(M, N, device=cuda) for _ in range(G)]
    aux = [_mk(M, N, cuda, 300 + g) for g in range(G)]
    cs_s, cs_g = [torch.zeros(N, device=cuda) for _ in range(G)], [torch.zeros(N, device=cuda) for _ in range(G)]
    with ops.grid_cap(cap):
        for g in range(G):
            kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn))
            if epi in (ops.EPI_BF16, ops.EPI_GELU, ops.EPI_RESID):
                kw["bias"] = biases[g]
            if epi == ops.EPI_GELU:
                kw["out2"] = extra_s[g]
            if epi == ops.EPI_RESID:
                kw["resid"] = resid[g]
            if epi == ops.EPI_MULAUX:
                kw.update(aux=aux[g], colsum=cs_s[g])
            if epi == ops.EPI_ATOMIC_F32:
                kw.update(splits=3, alpha=0.5)
            ops.gemm_bf16(As[g], Bs[g], epi, single[g], **kw)
        kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn))
        if epi in (ops.EPI_BF16, ops.EPI_GELU, ops.EPI_RESID):
            kw["biases"] = biases
        if epi == ops.EPI_GELU:
            kw["out2s"] = extra_g
        if epi == ops.EPI_RESID:
            kw["resids"] = resid
        if epi == ops.EPI_MULAUX:
            kw.update(auxs=aux, colsums=cs_g)
        if epi == ops.EPI_ATOMIC_F32:
            kw.update(splits=3, alpha=0.5)
        ops.gemm_bf16_grouped(As, Bs, epi, grouped, **kw)
    torch.cuda.synchronize()
    for g in range(G):
        if epi == ops.EPI_ATOMIC_F32:
            _close(grouped[g], single[g], 1e-5, 1e-5, f"group {g}")
        else:
            assert torch.equal(grouped[g], single[g]), f"group {g}"
        if epi == ops.EPI_GELU:
            assert torch.equal(extra_g[g], extra_s[g])
        if epi == ops.EPI_MULAUX:
            _close(cs_g[g], cs_s[g], 1e-5, 1e-4, f"colsum {g}")


@pytest.mark.parametrize("cap", [0, 3])
@pytest.mark.parametrize("B,N,H", [(5, 197, 3), (7, 64, 6), (4, 16, 2)])
def test_grouped_attention_equals_single_launches(B, N, H, cap, cuda):
    G = 3
    torch.manual_seed(5)
    qkvs = [torch.randn(B, N, 3, H, 64, device=cuda).to(torch.bfloat16) for _ in range(G)]
    douts = [(torch.randn(B, N, H * 64, device=cuda) * 0.1).to(torch.bfloat16) for _ in range(G)]
    with ops.grid_cap(cap):
        outs, lses = ops.attention_fwd_grouped(qkvs, B, N, H)
        dbg = [torch.zeros(3 * H * 64, device=cuda) for _ in range(G)]
        dqkvs = ops.attention_bwd_grouped(qkvs, outs, douts, lses, B, N, H, dbiases=dbg)
        for g in range(G):
            o, l = ops.attention_fwd(qkvs[g], B, N, H)
            db = torch.zeros(3 * H * 64, device=cuda)
            dq = ops.attention_bwd(qkvs[g], o, douts[g], l, B, N, H, dbias=db)
            assert torch.equal(o, outs[g]) and torch.equal(l, lses[g]), f"forward, group {g}"
            assert torch.equal(dq, dqkvs[g]), f"backward, group {g}"
            _close(dbg[g], db, 1e-5, 1e-5 * (B * N) ** 0.5, f"bias gradient, group {g}")
