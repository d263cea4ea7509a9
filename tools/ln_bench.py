"""Micro-benchmark of the LayerNorm kernels at the bench workload's shapes (img: 112*197 rows, txt: 112*64; d=384).
CUDA-graph replay over rotating buffers larger than L2.  Usage (GPU box): python tools/ln_bench.py"""
import sys

import torch

sys.path.insert(0, ".")
from fedcola_b200 import ops  # noqa: E402
from tools.attn_bench import time_graph  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    d = 384
    for rows in (112 * 197, 112 * 64):
        sets = 6
        x = [torch.randn(rows, d, device=dev) for _ in range(sets)]
        dy = [torch.randn(rows, d, device=dev).to(torch.bfloat16) for _ in range(sets)]
        dx = [torch.randn(rows, d, device=dev) for _ in range(sets)]
        dxs = [torch.empty(rows, d, device=dev, dtype=torch.bfloat16) for _ in range(sets)]
        g, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
        dg, db, cs = torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        scale = torch.ones(112, device=dev)
        _, mean, rstd = ops.layernorm_fwd(x[0], g, b, 1e-5, bf16_out=True)
        t_f = time_graph(lambda i: ops.layernorm_fwd(x[i], g, b, 1e-5, bf16_out=True), sets)
        t_b = time_graph(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g, dx[i], True, dxs=dxs[i], row_scale=scale,
                                                     rows_per_group=rows // 112, dgamma=dg, dbeta=db, dxs_colsum=cs), sets)
        by_f = rows * d * (4 + 2)
        by_b = rows * d * (2 + 4 + 4 + 4 + 2)
        print(f"rows={rows}: ln_fwd {t_f:6.1f} us ({by_f / t_f / 1e3:5.0f} GB/s)   ln_bwd {t_b:6.1f} us ({by_b / t_b / 1e3:5.0f} GB/s)",
              flush=True)
        # three clients per launch (a lockstep group): sets i, i+1, i+2 of the rotation
        G = 3
        gs, dgs, dbs, css = [g] * G, [dg] * G, [db] * G, [cs] * G
        t_g = time_graph(lambda i: ops.layernorm_bwd_grouped(
            [dy[(i + j) % sets] for j in range(G)], [x[(i + j) % sets] for j in range(G)], [mean] * G, [rstd] * G, gs,
            [dx[(i + j) % sets] for j in range(G)], True, dxs=[dxs[(i + j) % sets] for j in range(G)], row_scales=[scale] * G,
            rows_per_group=rows // 112, dgammas=dgs, dbetas=dbs, dxs_colsums=css), sets)
        print(f"rows=3x{rows} (grouped launch): ln_bwd {t_g:6.1f} us ({G * by_b / t_g / 1e3:5.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
