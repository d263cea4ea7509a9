"""BASELINE.json configs[4]: server aggregation microbenchmark — K clients x ViT-B-sized arenas, share_scope
modes, HBM GB/s (algorithmic bytes / CUDA-event time of the single fc_aggregate launch).
   python tools/agg_bench.py [--small]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fedcola_b200 import aggregation as agg  # noqa: E402
from fedcola_b200.arena import MatSpec  # noqa: E402

dev = torch.device("cuda:0")
small = "--small" in sys.argv
D, L, H = (384, 12, 6) if small else (768, 12, 12)
MODES = [("none", "dataset", False, False), ("attn", "modality", False, False), ("attn", "modality", True, True),
         ("blocks", "modality_exact", False, False), ("attn", "all", False, False), ("blocks", "all", False, False)]
KS = [8, 16, 32, 64] if small else [8, 16, 32, 64, 128]


def spec(kind, sp, sc, aux):
    kw = dict(embed_dim=D, depth=L, num_heads=H, max_text_len=64, with_aux=aux, aux_trained=aux, shared_param=sp,
              share_scope=sc)
    if kind == "img":
        return MatSpec(modalities=("img", None), num_classes=(100, None), tasks=("cls", None), **kw)
    if kind == "txt":
        return MatSpec(modalities=(None, "txt"), num_classes=(None, 4), tasks=(None, "cls"), **kw)
    return MatSpec(modalities=("img", "txt"), num_classes=(None, None), tasks=("rtv", "rtv"), vocab_size=7732, **kw)


print(f"# arena: d={D} L={L}; clients: 3/8 img, 3/8 txt, 1/4 img-txt; sizes randint(500,5000); times = median of 5 (L2 flushed)")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for sp, sc, comp, aux in MODES:
    specs = {k: spec(k, sp, sc, aux) for k in ("img", "txt", "pair")}
    meta = {"img": ("CIFAR100", "img", "cls"), "txt": ("AG_NEWS", "txt", "cls"), "pair": ("Flickr30k", "img+txt", "img+txt")}
    gl_ar = {k: torch.randn(s.total, device=dev) * 0.02 for k, s in specs.items()}
    names = []
    for s in specs.values():
        for k in s.keys():
            if k not in names:
                names.append(k)
    scope = agg.init_param_scope(names, sp, sc)
    flags = dict(args_modalities=["img", "txt", "img+txt", "img+txt"], share_scope_flag=sc, compensation=comp, with_aux=aux)
    for K in KS:
        g = torch.Generator().manual_seed(K)
        kinds = ["img"] * (3 * K // 8) + ["txt"] * (3 * K // 8)
        kinds += ["pair"] * (K - len(kinds))
        need = sum(specs[k].total for k in kinds) * 4 / 1e9
        if need > 120:
            print(f"{sp:6s} {sc:14s} comp={int(comp)} aux={int(aux)} K={K:4d}  skipped ({need:.0f} GB of client arenas)")
            continue
        cl = []
        for i, k in enumerate(kinds):
            a = gl_ar[k] + 0.02 * torch.randn(specs[k].total, device=dev)
            ds, m, t = meta[k]
            cl.append(agg.ClientCtx(i, ds, m, t, int(torch.randint(500, 5000, (1,), generator=g)), specs[k], a))
        outs = {k: torch.empty_like(v) for k, v in gl_ar.items()}
        gl = [agg.GlobalCtx(meta[k][0], meta[k][1], "rtv" if k == "pair" else "cls", 1, specs[k], gl_ar[k], outs[k])
              for k in ("img", "txt", "pair")]
        t0 = time.perf_counter()
        plan = agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags).to_device(dev)
        plan_ms = (time.perf_counter() - t0) * 1e3
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.launch()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[2] * 1e-3
        print(f"{sp:6s} {sc:14s} comp={int(comp)} aux={int(aux)} K={K:4d}  {plan.algorithmic_bytes/1e9:7.2f} GB  {t*1e3:8.3f} ms  "
              f"{plan.algorithmic_bytes/t/1e9:7.0f} GB/s  (plan {plan_ms:6.1f} ms host, {plan.n_jobs} jobs)")
        del cl, plan
        torch.cuda.empty_cache()
