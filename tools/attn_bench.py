"""Micro-benchmark of the attention kernels at the bench workload's shapes (img: N=197, txt: N=64; B=112, H=6).
Launches are captured in a CUDA graph (R rotating input sets > L2, so every launch reads HBM) to keep the host's
launch path out of the number.  Usage (GPU box): python tools/attn_bench.py"""
import sys

import torch

sys.path.insert(0, ".")
from fedcola_b200 import ops  # noqa: E402


def time_graph(fn, sets, reps=5):
    """fn(i) launches on input set i; returns median us per launch"""
    for i in range(sets):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(3):
            for i in range(sets):
                fn(i)
    n = 3 * sets
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    for (B, N, H) in [(112, 197, 6), (112, 64, 6), (112, 197, 12), (112, 40, 6), (112, 256, 6)]:
        per = B * N * 3 * H * 64 * 2
        sets = max(2, min(8, (200 << 20) // per + 1))
        qkv = [torch.randn(B, N, 3 * H * 64, device=dev).to(torch.bfloat16) for _ in range(sets)]
        dout = [(torch.randn(B, N, H * 64, device=dev) * 0.1).to(torch.bfloat16) for _ in range(sets)]
        ol = [ops.attention_fwd(q, B, N, H) for q in qkv]
        dbias = torch.zeros(3 * H * 64, device=dev)
        t_f = time_graph(lambda i: ops.attention_fwd(qkv[i], B, N, H), sets)
        t_b = time_graph(lambda i: ops.attention_bwd(qkv[i], ol[i][0], dout[i], ol[i][1], B, N, H, dbias=dbias), sets)
        fl = 4.0 * B * H * N * N * 64
        print(f"B={B} N={N} H={H}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TF/s, {1.33 * per / t_f / 1e3:5.0f} GB/s)   "
              f"bwd {t_b:7.1f} us ({2.5 * fl / t_b / 1e6:6.1f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
