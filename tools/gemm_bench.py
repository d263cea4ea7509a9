"""Micro-benchmark of the tcgen05 GEMM on the ViT-S client shapes (CUDA events, L2-flushing between launches).
   python tools/gemm_bench.py [--ncu]   (--ncu: one launch per shape, no timing loop)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fedcola_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
T, d = 112 * 197, 384
ncu = "--ncu" in sys.argv


def bf(r, c):
    return (torch.randn(r, c, device=dev) * 0.1).to(torch.bfloat16)


def run(name, fn, flops, bytes_):
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    if ncu:
        fn()
        torch.cuda.synchronize()
        return
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    print(f"{name:34s} {t*1e6:8.1f} us  {flops/t/1e12:7.1f} TFLOP/s  {bytes_/t/1e9:7.0f} GB/s (algorithmic)")


x = bf(T, d)
h = bf(T, 4 * d)
q3 = bf(T, 3 * d)
Wqkv, Wp, W1, W2 = bf(3 * d, d), bf(d, d), bf(4 * d, d), bf(d, 4 * d)
b3, b1, b4 = torch.randn(3 * d, device=dev), torch.randn(d, device=dev), torch.randn(4 * d, device=dev)
o3 = torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16)
o4a, o4b = torch.empty(T, 4 * d, device=dev, dtype=torch.bfloat16), torch.empty(T, 4 * d, device=dev, dtype=torch.bfloat16)
o1 = torch.empty(T, d, device=dev, dtype=torch.bfloat16)
res = torch.randn(T, d, device=dev)
of = torch.empty(T, d, device=dev)
gw3, gw4, gw1 = torch.zeros(3 * d, d, device=dev), torch.zeros(4 * d, d, device=dev), torch.zeros(d, 4 * d, device=dev)

run("fwd qkv   [T,384]x[1152,384] bf16", lambda: ops.gemm_bf16(x, Wqkv, ops.EPI_BF16, o3, bias=b3),
    2 * T * 3 * d * d, T * d * 2 + T * 3 * d * 2)
run("fwd proj  [T,384]x[384,384] resid", lambda: ops.gemm_bf16(x, Wp, ops.EPI_RESID, of, bias=b1, resid=res),
    2 * T * d * d, T * d * 2 + 2 * T * d * 4)
run("fwd fc1   [T,384]x[1536,384] gelu", lambda: ops.gemm_bf16(x, W1, ops.EPI_GELU, o4a, out2=o4b, bias=b4),
    2 * T * 4 * d * d, T * d * 2 + 2 * T * 4 * d * 2)
run("fwd fc2   [T,1536]x[384,1536] resid", lambda: ops.gemm_bf16(h, W2, ops.EPI_RESID, of, bias=b1, resid=res),
    2 * T * 4 * d * d, T * 4 * d * 2 + 2 * T * d * 4)
run("bwd dX fc2 [T,384]x[384,1536]mn mulaux+colsum", lambda: ops.gemm_bf16(x, W2, ops.EPI_MULAUX, o4a, b_mn=True, aux=o4b, colsum=b4),
    2 * T * 4 * d * d, T * d * 2 + 2 * T * 4 * d * 2)
run("bwd dX fc1 [T,1536]x[1536,384]mn", lambda: ops.gemm_bf16(h, W1, ops.EPI_BF16, o1, b_mn=True),
    2 * T * 4 * d * d, T * 4 * d * 2 + T * d * 2)
run("bwd dX qkv [T,1152]x[1152,384]mn", lambda: ops.gemm_bf16(q3, Wqkv, ops.EPI_BF16, o1, b_mn=True),
    2 * T * 3 * d * d, T * 3 * d * 2 + T * d * 2)
run("bwd dW qkv [T,1152]^T[T,384] split", lambda: ops.gemm_bf16(q3, x, ops.EPI_ATOMIC_F32, gw3, a_mn=True, b_mn=True, splits=0),
    2 * T * 3 * d * d, T * 3 * d * 2 + T * d * 2)
run("bwd dW fc1 [T,1536]^T[T,384] split", lambda: ops.gemm_bf16(h, x, ops.EPI_ATOMIC_F32, gw4, a_mn=True, b_mn=True, splits=0),
    2 * T * 4 * d * d, T * 4 * d * 2 + T * d * 2)
run("bwd dW fc2 [T,384]^T[T,1536] split", lambda: ops.gemm_bf16(x, h, ops.EPI_ATOMIC_F32, gw1, a_mn=True, b_mn=True, splits=0),
    2 * T * 4 * d * d, T * 4 * d * 2 + T * d * 2)
