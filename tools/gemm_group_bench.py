"""Does one grouped launch (G clients' same-layer GEMM) beat G single launches?  ViT-S layer shapes, CUDA events,
L2 flushed.   python tools/gemm_group_bench.py [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fedcola_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 112 * 197
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
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


G = 3
h = 4 * d
x = [bf(T, d) for _ in range(G)]
hh = [bf(T, h) for _ in range(G)]
q3 = [bf(T, 3 * d) for _ in range(G)]
Wqkv, Wp, W1, W2 = ([bf(*s) for _ in range(G)] for s in ((3 * d, d), (d, d), (h, d), (d, h)))
b3, b1, b4 = ([torch.randn(n, device=dev) for _ in range(G)] for n in (3 * d, d, h))
o3 = [torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16) for _ in range(G)]
o4a = [torch.empty(T, h, device=dev, dtype=torch.bfloat16) for _ in range(G)]
o4b = [torch.empty(T, h, device=dev, dtype=torch.bfloat16) for _ in range(G)]
o1 = [torch.empty(T, d, device=dev, dtype=torch.bfloat16) for _ in range(G)]
res = [torch.randn(T, d, device=dev) for _ in range(G)]
of = [torch.empty(T, d, device=dev) for _ in range(G)]
gw3, gw4, gw1 = ([torch.zeros(*s, device=dev) for _ in range(G)] for s in ((3 * d, d), (h, d), (d, h)))

cases = [
    ("fwd qkv bf16", 2 * T * 3 * d * d, lambda g: ops.gemm_bf16_grouped(x[:g], Wqkv[:g], ops.EPI_BF16, o3[:g], biases=b3[:g])),
    ("fwd proj resid", 2 * T * d * d, lambda g: ops.gemm_bf16_grouped(x[:g], Wp[:g], ops.EPI_RESID, of[:g], biases=b1[:g], resids=res[:g])),
    ("fwd fc1 gelu", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(x[:g], W1[:g], ops.EPI_GELU, o4a[:g], out2s=o4b[:g], biases=b4[:g])),
    ("fwd fc2 resid", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(hh[:g], W2[:g], ops.EPI_RESID, of[:g], biases=b1[:g], resids=res[:g])),
    ("dX fc2 mulaux", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(x[:g], W2[:g], ops.EPI_MULAUX, o4a[:g], b_mn=True, auxs=o4b[:g], colsums=b4[:g])),
    ("dX fc1 bf16", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(hh[:g], W1[:g], ops.EPI_BF16, o1[:g], b_mn=True)),
    ("dX qkv bf16", 2 * T * 3 * d * d, lambda g: ops.gemm_bf16_grouped(q3[:g], Wqkv[:g], ops.EPI_BF16, o1[:g], b_mn=True)),
    ("dW qkv split", 2 * T * 3 * d * d, lambda g: ops.gemm_bf16_grouped(q3[:g], x[:g], ops.EPI_ATOMIC_F32, gw3[:g], a_mn=True, b_mn=True, splits=0)),
    ("dW fc1 split", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(hh[:g], x[:g], ops.EPI_ATOMIC_F32, gw4[:g], a_mn=True, b_mn=True, splits=0)),
    ("dW fc2 split", 2 * T * h * d, lambda g: ops.gemm_bf16_grouped(x[:g], hh[:g], ops.EPI_ATOMIC_F32, gw1[:g], a_mn=True, b_mn=True, splits=0)),
]
print(f"# T={T} d={d} PAIR={os.environ.get('FC_GEMM_PAIR', '0')}")
tot = [0.0, 0.0, 0.0]
for name, fl, fn in cases:
    ts = [timeit(lambda g=g: fn(g)) for g in (1, 2, 3)]
    for i in range(3):
        tot[i] += ts[i] / (i + 1)
    print(f"{name:16s}  G=1 {ts[0]:6.1f} us ({fl/ts[0]/1e6:6.0f} TF/s)   G=2 {ts[1]:6.1f} us = {ts[1]/2:5.1f}/client ({2*fl/ts[1]/1e6:6.0f})   "
          f"G=3 {ts[2]:6.1f} us = {ts[2]/3:5.1f}/client ({3*fl/ts[2]/1e6:6.0f} TF/s)", flush=True)
print(f"sum per client: G=1 {tot[0]:.1f} us, G=2 {tot[1]:.1f} us, G=3 {tot[2]:.1f} us")
