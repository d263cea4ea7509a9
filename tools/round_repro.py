"""A few federated rounds of a bench configuration outside bench.py (sanitizer / debugger target).
Usage (GPU box): python tools/round_repro.py [--resident host|device] [--rounds 2] [--threads 3] [--config vits-flickr] [--n 896]
e.g.  compute-sanitizer --tool memcheck python tools/round_repro.py --resident host --rounds 1 --n 224"""
import argparse
import os
import random
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--resident", default="host")
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--threads", type=int, default=3)
ap.add_argument("--client-group", type=int, default=3)
ap.add_argument("--config", default="vits-flickr")
ap.add_argument("--n", type=int, default=0, help="samples per client (0 = the configuration's)")
a = ap.parse_args()
sys.argv = sys.argv[:1]                     # bench.py reads the command line when imported
import bench  # noqa: E402
from fedcola_b200.harness import make_client_datasets  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = dict(bench.CONFIGS[a.config])
if a.n:
    cfg["n"] = a.n
counts = cfg["per_gpu"] or cfg["total"]
args = bench.workload_args(cfg, counts, data_resident=a.resident, server_device=str(dev), num_thread=a.threads,
                           client_devices=[str(dev)], placement="reference", precision="bf16", client_group=a.client_group)
random.seed(args.seed)
torch.manual_seed(args.seed)
cds = make_client_datasets(bench.client_specs(cfg, counts), seq_len=bench.SEQ, share=True)
cls = import_module(f"fedcola_b200.server.{args.algorithm}server").__dict__[f"{args.algorithm.title()}Server"]
server = cls(args=args, writer=bench.NullWriter(), server_dataset=(None, {}), client_datasets=cds, model_str=cfg["model"])
for r in range(a.rounds):
    server.round += 1
    ids = server.update()
    torch.cuda.synchronize()
    print(f"round {server.round}: clients {ids} ok, local training {server.phase_ms['local_training']:.1f} ms", flush=True)
