"""GPU: run-to-run determinism of the tcgen05 kernels.  out / lse / dqkv of the attention kernels and the non-atomic GEMM
epilogues have no atomics on their path: repeated launches on the same inputs must be BIT-identical — a mismatch is a
race inside the kernel (mbarrier protocol, staging-tile reuse, TMEM hand-over), which tolerance-based parity tests can
miss.  The atomically accumulated outputs (bias gradients) only have to agree to fp32 reordering noise."""
import pytest
import torch

from fedcola_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,H,G", [(6, 16, 2, 1), (12, 33, 6, 3), (112, 64, 6, 3), (64, 100, 6, 2), (112, 197, 6, 1),
                                     (112, 197, 6, 3), (16, 256, 6, 1)])
def test_attention_is_bit_reproducible(B, N, H, G, cuda):
    torch.manual_seed(B * 1000 + N)
    qkv = [torch.randn(B, N, 3 * H * 64, device=cuda).to(torch.bfloat16) for _ in range(G)]
    dout = [(torch.randn(B, N, H * 64, device=cuda) * 0.1).to(torch.bfloat16) for _ in range(G)]
    ref_f = ref_b = ref_db = None
    for rep in range(6):
        if G == 1:
            o, l = ops.attention_fwd(qkv[0], B, N, H)
            outs, lses = [o], [l]
        else:
            outs, lses = ops.attention_fwd_grouped(qkv, B, N, H)
        torch.cuda.synchronize()
        cur = [t.clone() for t in outs] + [t.clone() for t in lses]
        if ref_f is None:
            ref_f = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref_f, cur)), f"forward differs in repetition {rep}"
        dbs = [torch.zeros(3 * H * 64, device=cuda) for _ in range(G)]
        if G == 1:
            dq = [ops.attention_bwd(qkv[0], ref_f[0], dout[0], ref_f[G], B, N, H, dbias=dbs[0])]
        else:
            dq = ops.attention_bwd_grouped(qkv, ref_f[:G], dout, ref_f[G:], B, N, H, dbiases=dbs)
        torch.cuda.synchronize()
        if ref_b is None:
            ref_b, ref_db = [t.clone() for t in dq], [t.clone() for t in dbs]
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref_b, dq)), f"backward differs in repetition {rep}"
            for a, b in zip(ref_db, dbs):                     # atomics: order noise only
                assert (a - b).abs().max().item() <= 1e-4 * a.abs().max().item() + 1e-30


@pytest.mark.parametrize("M,N,K,G,bmn", [(240, 96, 32, 3, False), (720, 384, 384, 3, False), (22064, 1152, 384, 1, False),
                                         (22064, 384, 1536, 1, True), (7168, 1536, 384, 3, False), (1000, 200, 72, 2, False)])
def test_gemm_is_bit_reproducible(M, N, K, G, bmn, cuda):
    torch.manual_seed(M + N)
    As = [(torch.randn(M, K, device=cuda) * 0.5).to(torch.bfloat16) for _ in range(G)]
    Bs = [(torch.randn((K, N) if bmn else (N, K), device=cuda) * 0.5).to(torch.bfloat16) for _ in range(G)]
    bias = [torch.randn(N, device=cuda) for _ in range(G)]
    resid = [torch.randn(M, N, device=cuda) for _ in range(G)]
    for epi in ((ops.EPI_BF16,) if bmn else (ops.EPI_BF16, ops.EPI_GELU, ops.EPI_RESID, ops.EPI_F32)):
        ref = None
        for rep in range(5):
            f32 = epi in (ops.EPI_RESID, ops.EPI_F32)
            outs = [torch.empty(M, N, device=cuda, dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(G)]
            out2 = [torch.empty(M, N, device=cuda, dtype=torch.bfloat16) for _ in range(G)] if epi == ops.EPI_GELU else None
            ops.gemm_bf16_grouped(As, Bs, epi, outs, b_mn=bmn, out2s=out2, biases=bias,
                                  resids=resid if epi == ops.EPI_RESID else None)
            torch.cuda.synchronize()
            cur = outs + (out2 or [])
            if ref is None:
                ref = [t.clone() for t in cur]
            else:
                assert all(torch.equal(a, b) for a, b in zip(ref, cur)), f"epilogue {epi} differs in repetition {rep}"
