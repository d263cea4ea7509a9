"""Fixed overhead vs steady-state rate of the GEMM kernel: time(M) for M = T/2 .. 4T at the ViT-S layer shapes,
least-squares fit  t = F + M/rate.   python tools/gemm_scaling.py   (env: FC_GEMM_PAIR, FC_GEMM_BN, FC_GEMM_DEBUG)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fedcola_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
T, d = 112 * 197, 384
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def bf(r, c):
    return (torch.randn(r, c, device=dev) * 0.1).to(torch.bfloat16)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def fit(xs, ys):
    n = len(xs)
    mx, my = sum(xs) / n, sum(ys) / n
    b = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sum((x - mx) ** 2 for x in xs)
    return my - b * mx, b


cases = {
    "fwd qkv  N=1152 K=384 bf16": lambda M: (lambda x=bf(M, d), w=bf(3 * d, d), o=torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16):
                                             ops.gemm_bf16(x, w, ops.EPI_BF16, o), 2.0 * 3 * d * d),
    "fwd fc1  N=1536 K=384 gelu": lambda M: (lambda x=bf(M, d), w=bf(4 * d, d), o=torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16),
                                             o2=torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16): ops.gemm_bf16(x, w, ops.EPI_GELU, o, out2=o2),
                                             2.0 * 4 * d * d),
    "fwd fc2  N=384 K=1536 resid": lambda M: (lambda x=bf(M, 4 * d), w=bf(d, 4 * d), o=torch.empty(M, d, device=dev), r=torch.randn(M, d, device=dev):
                                              ops.gemm_bf16(x, w, ops.EPI_RESID, o, resid=r), 2.0 * 4 * d * d),
    "dX fc1   N=384 K=1536 bf16 (B mn)": lambda M: (lambda x=bf(M, 4 * d), w=bf(4 * d, d), o=torch.empty(M, d, device=dev, dtype=torch.bfloat16):
                                                    ops.gemm_bf16(x, w, ops.EPI_BF16, o, b_mn=True), 2.0 * 4 * d * d),
    "dW fc1   [1536x384] K=M split": lambda M: (lambda a=bf(M, 4 * d), b=bf(M, d), o=torch.zeros(4 * d, d, device=dev):
                                                ops.gemm_bf16(a, b, ops.EPI_ATOMIC_F32, o, a_mn=True, b_mn=True, splits=0), 2.0 * 4 * d * d),
}
print(f"# PAIR={os.environ.get('FC_GEMM_PAIR', '0')} BN={os.environ.get('FC_GEMM_BN', 'auto')} DEBUG={os.environ.get('FC_GEMM_DEBUG', '0')}")
for name, mk in cases.items():
    Ms = [T // 2, T, 2 * T, 4 * T]
    ts = []
    for M in Ms:
        fn, fpr = mk(M)
        ts.append(timeit(fn))
        del fn
        torch.cuda.empty_cache()
    F, slope = fit(Ms, ts)
    rate = fpr / slope / 1e6          # flops per row / (us per row) -> TFLOP/s
    print(f"{name:36s} t(us) = " + " ".join(f"{t:7.1f}" for t in ts) + f"   fixed {F:6.1f} us, steady {rate:7.1f} TFLOP/s", flush=True)
