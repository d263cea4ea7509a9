"""Run-to-run determinism of the attention kernels: out / lse / dqkv have no atomics on their path, so repeated launches
on the same inputs must be bit-identical (a mismatch = a race inside the kernel).  python tools/determinism_check.py"""
import sys

import torch

sys.path.insert(0, ".")
from fedcola_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bad = 0
for (B, N, H, G) in [(6, 16, 2, 1), (6, 40, 2, 3), (12, 33, 6, 3), (112, 64, 6, 1), (112, 64, 6, 3), (64, 100, 6, 2), (112, 197, 6, 1),
                     (112, 197, 6, 3), (32, 256, 6, 1), (500, 40, 6, 2)]:
    torch.manual_seed(B * 1000 + N)
    qkv = [torch.randn(B, N, 3 * H * 64, device=dev).to(torch.bfloat16) for _ in range(G)]
    dout = [(torch.randn(B, N, H * 64, device=dev) * 0.1).to(torch.bfloat16) for _ in range(G)]
    ref_f = ref_b = None
    nf = nb = 0
    for rep in range(25):
        if G == 1:
            o, l = ops.attention_fwd(qkv[0], B, N, H)
            outs, lses = [o], [l]
        else:
            outs, lses = ops.attention_fwd_grouped(qkv, B, N, H)
        torch.cuda.synchronize()
        cur = [t.clone() for t in outs] + [t.clone() for t in lses]
        if ref_f is None:
            ref_f = cur
        elif any(not torch.equal(a, b) for a, b in zip(ref_f, cur)):
            nf += 1
        dbs = [torch.zeros(3 * H * 64, device=dev) for _ in range(G)]
        if G == 1:
            dq = [ops.attention_bwd(qkv[0], ref_f[0], dout[0], ref_f[G], B, N, H, dbias=dbs[0])]
        else:
            dq = ops.attention_bwd_grouped(qkv, ref_f[:G], dout, ref_f[G:], B, N, H, dbiases=dbs)
        torch.cuda.synchronize()
        cur = [t.clone() for t in dq]
        if ref_b is None:
            ref_b, ref_db = cur, [t.clone() for t in dbs]
        else:
            if any(not torch.equal(a, b) for a, b in zip(ref_b, cur)):
                nb += 1
            worst = max(((a - b).abs().max() / (a.abs().max() + 1e-20)).item() for a, b in zip(ref_db, dbs))
            if worst > 1e-4:
                nb += 1000
    print(f"B={B} N={N} H={H} groups={G}: forward mismatches {nf}/24, backward mismatches {nb}/24", flush=True)
    bad += nf + nb
print("DETERMINISTIC" if bad == 0 else f"NON-DETERMINISTIC ({bad})")

# ---- GEMM: the non-atomic epilogues must be bit-identical run to run too
bad = 0
for (M, N, K, G, amn, bmn) in [(240, 96, 32, 3, False, False), (240, 128, 32, 3, False, True), (720, 384, 384, 3, False, False),
                               (22064, 1152, 384, 1, False, False), (66192, 384, 1536, 1, False, True), (7168, 1536, 384, 3, False, False),
                               (1000, 200, 72, 2, False, False)]:
    torch.manual_seed(M + N)
    As = [(torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16) for _ in range(G)]
    Bs = [(torch.randn((K, N) if bmn else (N, K), device=dev) * 0.5).to(torch.bfloat16) for _ in range(G)]
    bias = [torch.randn(N, device=dev) for _ in range(G)]
    resid = [torch.randn(M, N, device=dev) for _ in range(G)]
    for epi in ((0,) if bmn else (0, 1, 2, 4)):
        ref, n = None, 0
        for rep in range(20):
            outs = [torch.empty(M, N, device=dev, dtype=torch.float32 if epi in (2, 4) else torch.bfloat16) for _ in range(G)]
            out2 = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(G)] if epi == 1 else None
            ops.gemm_bf16_grouped(As, Bs, epi, outs, b_mn=bmn, out2s=out2, biases=bias, resids=resid if epi == 2 else None)
            torch.cuda.synchronize()
            cur = outs + (out2 or [])
            if ref is None:
                ref = [t.clone() for t in cur]
            elif any(not torch.equal(a, b) for a, b in zip(ref, cur)):
                n += 1
        if n:
            print(f"GEMM M={M} N={N} K={K} groups={G} epi={epi}: {n}/19 mismatching runs", flush=True)
        bad += n
print("GEMM DETERMINISTIC" if bad == 0 else f"GEMM NON-DETERMINISTIC ({bad})")
