"""One LayerNorm backward launch at the bench's image-group shape (for ncu).  python tools/ln_once.py [rows] [d]"""
import sys

import torch

sys.path.insert(0, ".")
from fedcola_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 112 * 197
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
dev = torch.device("cuda:0")
x = torch.randn(rows, d, device=dev)
dy = torch.randn(rows, d, device=dev).to(torch.bfloat16)
dx = torch.randn(rows, d, device=dev)
dxs = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
g, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
dg, db, cs = torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
scale = torch.ones(112, device=dev)
_, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, bf16_out=True)
for _ in range(3):
    ops.layernorm_bwd(dy, x, mean, rstd, g, dx, True, dxs=dxs, row_scale=scale, rows_per_group=max(rows // 112, 1), dgamma=dg,
                      dbeta=db, dxs_colsum=cs)
torch.cuda.synchronize()
