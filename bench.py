"""Benchmark of the FedCola round hot path (BASELINE.json metric: local-train samples/s per round @N B200;
aggregation HBM GB/s).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one full federated round of BASELINE.json configs[1] — FedCola (shared_param=attn,
share_scope=modality, --compensation --with_aux --aux_trained), ViT-S/16-sized ModalityAgnosticTransformer,
Flickr30k-shaped synthetic clients: per GPU 12 img + 12 txt + 8 img-txt clients, C=0.25 -> 3+3+2 sampled per
round, 896 samples each, B=112, E=1, seq_len 64, AdamW lr 1e-4, DropPath 0.1 — through the drop-in
`FedavgServer.update()`: sampling, download, local training of every sampled client, aggregation of the three
global models, aux refresh.  N GPUs = N x the clients (weak scaling), sharded by the server, NCCL all-reduce
of the closed-form partial aggregates.

`value`  : samples/s with every client's data resident in HBM (device-timed, max over ranks).
`e2e`    : same rounds with the data in pinned HOST memory: every step's batch is copied host->device inside
           the timed region and the per-epoch loss/acc statistics are read back device->host.
`--impl reference`: the reference's CPU path (oracle port, torch fp32 on the host cores) on a bounded sample of
the same workload.  Prints ONE JSON line.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOPS_PER_SAMPLE = {"img": 27.59e9, "txt": 8.38e9, "img+txt": 35.97e9}     # ViT-S fwd+bwd, BASELINE.md §3
MODEL = "mome_small_patch16"
B, N_PER_CLIENT, SEQ = 112, 896, 64


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def workload_args(n_gpus, **over):
    from fedcola_b200.harness import make_args
    kw = dict(model_name=MODEL, datasets=["CIFAR100", "AG_NEWS", "Flickr30k", "Coco"],
              modalities=["img", "txt", "img+txt", "img+txt"], shared_param="attn", share_scope="modality",
              compensation=True, with_aux=True, aux_trained=True, Ks=[12 * n_gpus, 12 * n_gpus, 8 * n_gpus],
              K=32 * n_gpus, Cs=[0.25], equal_sampled=True, B=B, E=1, optimizer="AdamW", lr=1e-4, lr_decay=0.99,
              lr_decay_step=1, seq_len=SEQ, dropout=0.1, no_shuffle=False, num_thread=1, seed=1,
              out_modality_scales=[1, 1, 1, 1], droppath_rng="fused", data_resident="device")
    kw.update(over)
    return make_args(**kw)


def client_specs(n_gpus, n=N_PER_CLIENT):
    return [("CIFAR100", n, 1)] * (12 * n_gpus) + [("AG_NEWS", n, 2)] * (12 * n_gpus) + [("Flickr30k", n, 3)] * (8 * n_gpus)


class NullWriter:
    def log(self, *a, **k):
        pass

    def finish(self):
        pass


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None
        self.t_begin = None        # only samples taken after mark_begin() (the timed region) are reported

    def mark_begin(self):
        self.t_begin = time.time()

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.t_begin is not None:
                    self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


_T0 = time.time()


def _log(msg):
    if int(os.environ.get("RANK", 0)) == 0:
        sys.stderr.write(f"[bench +{time.time() - _T0:6.1f}s] {msg}\n")
        sys.stderr.flush()


PREWARM_THREADED = 3


def _agg_traffic():
    """DRAM bytes (read + write) of the round's aggregation launch from the committed ncu capture, or None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_aggregate_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


def _gemm_traffic():
    """DRAM bytes per GEMM launch from the committed ncu capture (profiles/r1_gemm_traffic.json), or (None, why)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_gemm_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return round(t["avg_dram_bytes_per_launch"]), t["source"] + " — image-client shapes, cold-cache replays"
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture committed"


def run_ours(a):
    import torch.distributed as dist
    from fedcola_b200 import _lib
    from fedcola_b200.harness import make_client_datasets
    from fedcola_b200.server.fedavgserver import FedavgServer
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // max(world, 1)))    # torchrun pins OMP_NUM_THREADS=1
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _log("process group ready")
    n_gpus = max(world, 1)
    L = _lib.lib()
    L.fc_launch_count.restype = __import__("ctypes").c_ulonglong
    L.fc_gemm_profile_collect.restype = __import__("ctypes").c_longlong
    pk = peaks()

    def make_server(resident):
        args = workload_args(n_gpus, data_resident=resident, server_device=str(dev), num_thread=a.threads)
        random.seed(args.seed)
        torch.manual_seed(args.seed)
        cds = make_client_datasets(client_specs(n_gpus), seq_len=SEQ, share=True)
        return FedavgServer(args=args, writer=NullWriter(), server_dataset=(None, {}), client_datasets=cds,
                            model_str=MODEL), args

    def timed_rounds(server, steps, warmup, sampler=None):
        per, agg_ms, agg_bytes = [], [], []
        samples = 0
        phases = {"local_training": 0.0, "aggregation_and_refresh": 0.0}
        for it in range(warmup + steps):
            server.round += 1
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if it == warmup and sampler is not None:
                sampler.mark_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            ids = server.update()
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            if it >= warmup:
                # device time of the round; the round also has host work (sampling, planning), so take the
                # larger of the CUDA-event span and the wall clock between the two synchronisations
                per.append(max(e0.elapsed_time(e1), wall))
                la = server.last_aggregation
                agg_ms.append(la["events"][0].elapsed_time(la["events"][1]))
                agg_bytes.append(la["bytes"])
                samples += sum(server.args.E * len(server.clients[i]) for i in ids)
                for k in phases:
                    phases[k] += server.phase_ms[k] / steps
        timed_rounds.phases = phases
        timed_rounds.per_round = [round(x, 1) for x in per]
        t = torch.tensor([sum(per)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), samples, agg_ms, agg_bytes

    if a.profile:      # ncu --profile-from-start off: one warm round, then ONE profiled round
        server, args = make_server("device")
        server.round += 1
        server.update()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        server.round += 1
        server.update()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- kernel-side number: client data resident in HBM ---------------------------------------------
    server, args = make_server("device")
    _log("server built (HBM-resident data)")
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("FC_BENCH_NO_CLOCKS")) else None
    if sampler is not None:
        sampler.start()        # spawn nvidia-smi now (forking this process mid-run costs ~0.3 s of host time)
    l0 = L.fc_launch_count()
    # Client worker threads run on their own CUDA streams: torch's caching allocator keeps one pool per stream, so
    # a few extra untimed rounds let every (thread, stream) pool reach its steady size before the W warm-up rounds.
    warm = a.warmup + (PREWARM_THREADED if a.threads > 1 else 0)
    total_ms, samples, agg_ms, agg_bytes = timed_rounds(server, a.steps, warm, sampler)
    launches = (L.fc_launch_count() - l0)
    _log("timed rounds done")
    clocks = sampler.finish() if sampler is not None else None
    launches_timed = int(launches * a.steps / (a.steps + warm))
    value = samples / (total_ms / 1e3)
    phases = dict(timed_rounds.phases)
    per_round_ms = list(timed_rounds.per_round)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): one extra round with per-launch CUDA events ----
    # (one client at a time: with several client streams in flight an event pair around a launch would also time
    #  the other streams' kernels it waits behind)
    L.fc_gemm_profile(1)
    server.round += 1
    threads, server.args.num_thread = server.args.num_thread, 1
    ids = server.update()
    server.args.num_thread = threads
    torch.cuda.synchronize()
    import ctypes
    ms, fl = ctypes.c_double(0), ctypes.c_double(0)
    n_gemm = L.fc_gemm_profile_collect(ctypes.byref(ms), ctypes.byref(fl))
    L.fc_gemm_profile(0)
    gemm_tflops = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
    round_flops = sum(FLOPS_PER_SAMPLE[server.clients[i].modality] * len(server.clients[i]) for i in ids
                      if server._owner.get(i, 0) == rank)
    agg_gbs = [b / (m * 1e-3) / 1e9 for b, m in zip(agg_bytes, agg_ms) if m > 0]
    agg_best = sorted(agg_gbs)[len(agg_gbs) // 2] if agg_gbs else 0.0

    # ---- end to end: same rounds, client data in pinned host memory (H2D per batch, D2H stats per epoch) ----
    del server
    torch.cuda.empty_cache()
    _log("GEMM profile round done")
    server2, _ = make_server("host")
    _log("server built (host-resident data)")
    e2e_ms, e2e_samples, _, _ = timed_rounds(server2, a.steps, max(warm, 1))
    per_sample = {"img": 3 * 224 * 224 * 4 + 8, "txt": SEQ * 8 + 8, "img+txt": 3 * 224 * 224 * 4 + SEQ * 8}
    h2d = sum(per_sample[m] * N_PER_CLIENT * c for m, c in (("img", 3), ("txt", 3), ("img+txt", 2))) * n_gpus
    d2h = 16 * 8 * n_gpus
    e2e_value = e2e_samples / (e2e_ms / 1e3)
    _log("e2e rounds done")

    if rank == 0:
        out = {
            "metric": "local_train_samples_per_s_per_round", "value": round(value, 2), "unit": "samples/s",
            "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(total_ms / a.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "FedCola round: ViT-S/16-sized MAT, shared_param=attn share_scope=modality "
                                   "+compensation +with_aux +aux_trained; per GPU 12 img + 12 txt + 8 img-txt "
                                   "Flickr30k-shaped synthetic clients, C=0.25 -> 3+3+2 sampled, 896 samples/client, "
                                   "B=112, E=1, seq_len 64, AdamW lr 1e-4, drop_path 0.1; step = one server.update()",
                       "samples_per_round": samples // a.steps, "precision": "bf16 operands / fp32 accumulate, fp32 "
                       "master weights+optimizer+aggregation", "l2": "inputs larger than L2 (GBs of activations and "
                       "parameters per round)", "parallelism": f"clients sharded over {n_gpus} GPU(s); {a.threads} client worker thread(s)/GPU "
                                      f"(args.num_thread), each on its own CUDA stream"},
            "e2e": {"value": round(e2e_value, 2), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_timed,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": round(gemm_tflops, 2),
                         "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(gemm_tflops / pk["tf_sustained"], 4),
                         "peak_source": pk["src"] + " sustained bf16", "traffic": _gemm_traffic()[0], "traffic_source": _gemm_traffic()[1],
                         "launches_per_round": int(n_gemm), "avg_launch_us": round(ms.value * 1e3 / max(n_gemm, 1), 2),
                         "round_model_tflops": round(round_flops * n_gpus / (total_ms / a.steps * 1e-3) / 1e12 / n_gpus, 2)},
            "phase_ms_host_clock": {k: round(v, 2) for k, v in phases.items()},
            "per_round_ms": per_round_ms, "e2e_per_round_ms": list(timed_rounds.per_round),
            "train_phase_samples_per_s": round(samples / a.steps / (phases["local_training"] * 1e-3), 2),
            "aggregation": {"gbs": round(agg_best, 1), "frac_of_measured_hbm": round(agg_best / pk["hbm"], 4),
                            "bytes_per_round": int(agg_bytes[0]) if agg_bytes else 0,
                            "ms": round(sorted(agg_ms)[len(agg_ms) // 2], 4) if agg_ms else None,
                            "traffic": _agg_traffic()},
        }
        if a.cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def oracle_round_server(n_img=32, n_txt=32, n_pair=32, batch=32):
    """Oracle port of the same round on a bounded sample: one client of each kind."""
    from fedcola_b200.arena import MatSpec
    from fedcola_b200.harness import make_client_datasets
    from oracle.round_oracle import OracleServer
    import numpy as np
    args = workload_args(1, Ks=[1, 1, 1], K=3, Cs=[1.0], B=batch, dropout=0.0)
    cds = make_client_datasets([("CIFAR100", n_img, 1), ("AG_NEWS", n_txt, 2), ("Flickr30k", n_pair, 3)], seq_len=SEQ)
    specs, init = {}, {}
    rng = np.random.RandomState(0)
    for ds, mods, ncls, tasks, vocab in (("CIFAR100", ("img", None), (100, None), ("cls", None), 30522),
                                         ("AG_NEWS", (None, "txt"), (None, 4), (None, "cls"), 30522),
                                         ("Flickr30k", ("img", "txt"), (None, None), ("rtv", "rtv"), 7732)):
        sp = MatSpec(embed_dim=384, depth=12, num_heads=6, modalities=mods, num_classes=ncls, tasks=tasks,
                     vocab_size=vocab, max_text_len=SEQ, with_aux=True, aux_trained=True, shared_param="attn",
                     share_scope="modality")
        specs[ds] = sp
        st = {}
        for s in sp.unique_segments():
            v = (rng.standard_normal(s.numel) * 0.02).astype(np.float32).reshape(s.shape)
            if "norm" in s.key.lower() and s.key.endswith("weight"):
                v = v + 1.0
            if s.key.endswith("cross_modal_scale"):
                v = np.zeros(s.shape, np.float32)
            st[s.key] = v
        init[ds] = st
    random.seed(1)
    torch.manual_seed(1)
    return OracleServer(args, cds, specs, init)


def cpu_baseline_sample():
    torch.set_num_threads(os.cpu_count() or 1)
    srv = oracle_round_server(16, 16, 16, 16)
    srv.round = 1
    t0 = time.perf_counter()
    srv.update()
    dt = time.perf_counter() - t0
    return {"value": round(srv.timing["samples"] / dt, 3), "unit": "samples/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": "one full round of the same FedCola/ViT-S workload on 3 clients (1 img, 1 txt, "
            "1 img-txt) x 16 samples, B=16, oracle port (torch fp32 CPU)",
            "train_samples_per_s": round(srv.timing["samples"] / srv.timing["train_s"], 3),
            "aggregation_gbs": round(srv.timing["agg_bytes"] / srv.timing["agg_s"] / 1e9, 3)}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n = 16
    srv = oracle_round_server(n, n, n, n)
    total, samples = 0.0, 0
    for it in range(a.warmup + a.steps):
        srv.round += 1
        t0 = time.perf_counter()
        srv.update()
        dt = time.perf_counter() - t0
        if it >= a.warmup:
            total += dt
            samples += srv.timing["samples"]
    v = samples / total
    sample = f"each step = one full round on 3 clients (1 img, 1 txt, 1 img-txt) x {n} samples, B={n}, oracle port of the reference (torch fp32, CPU)"
    print(json.dumps({
        "impl": "reference", "metric": "local_train_samples_per_s_per_round", "value": round(v, 3), "unit": "samples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(total / a.steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "FedCola round, ViT-S/16-sized MAT (same config as the GPU arm), bounded sample", "sample": sample},
        "cpu_baseline": {"value": round(v, 3), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--threads", type=int, default=3, help="client worker threads per GPU (args.num_thread)")
    ap.add_argument("--profile", action="store_true", help="one warm + one cudaProfiler-bracketed round (for ncu)")
    a = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything libraries print while running (NCCL's version
    # banner, warnings) is routed to stderr at file-descriptor level, and fd 1 is handed back for the final print.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_print = print

    def emit(*args, **kw):
        sys.stdout.flush()
        os.dup2(saved, 1)
        real_print(*args, **kw)
        sys.stdout.flush()
        os.dup2(2, 1)

    globals()["print"] = emit            # run_ours / run_reference print only their JSON line
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
