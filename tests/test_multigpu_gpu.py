"""GPU, >= 2 devices (skipped otherwise; run with `gpurun --gpus 2`): hardware evidence for the two multi-GPU modes.

 * single process, clients on `cuda:(i % ngpu)` — the reference's own mode (fedavgserver.py:310-311, 560-577): the
   aggregation kernel on the server GPU reads the other GPU's client arenas in place over NVLink peer access and the
   sequential lerp stays BIT-exact (reference golden hashes);
 * one process per GPU (torchrun, NCCL): closed-form partial sums + one all-reduce of the compact staging buffer must
   equal the 1-GPU sequential-lerp round to <= 1e-6 (norm-wise relative; SURVEY §4 "distributed" row), with the same
   sampled ids and the same logged loss on every rank."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as H
from fedcola_b200 import aggregation as agg

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "agg_hashes.json")


@pytest.fixture(scope="module")
def two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    return torch.device("cuda:0"), torch.device("cuda:1")


@pytest.mark.parametrize("case", ["fedcola_attn_modality_comp_aux", "attn_modality_scaled", "blocks_all_comp"])
def test_peer_read_aggregation_is_bit_exact(case, two_gpus):
    from fedcola_b200 import _lib
    d0, d1 = two_gpus
    assert _lib.lib().fc_enable_peer_access(0, 1) == 0, _lib.lib().fc_last_error()
    with open(GOLDEN) as f:
        golden = json.load(f)[case]
    gl, cl, scope, flags = H.build_agg_case(case, device=d0)
    for c in cl[::2]:
        c.arena = c.arena.to(d1)                  # every other client was "trained" on the second GPU
    torch.cuda.synchronize(d1)
    with torch.cuda.device(d0):
        agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags).to_device(d0).launch()
    torch.cuda.synchronize(d0)
    for g in gl:
        got = H.state_dict_of(g.spec, g.arena_out.cpu().numpy())
        for k, h in golden[g.dataset].items():
            assert H.sha(got[k]) == h, (case, g.dataset, k)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("case", ["fedcola", "fediot"])
def test_single_process_two_gpu_round_equals_one_gpu_round(case, two_gpus):
    from test_round_gpu import our_round
    d0, _ = two_gpus
    one, ids1, datasets, _ = our_round(case, d0, client_devices=["cuda:0"])
    two, ids2, _, _ = our_round(case, d0, client_devices=["cuda:0", "cuda:1"], num_thread=2)
    assert ids1 == ids2
    assert {c.device for c in two.clients} == {"cuda:0", "cuda:1"}
    for ds in datasets:
        # training is not bit-reproducible run to run (fp32 atomics in the split-K / bias-gradient reductions); the
        # aggregation itself is the same kernel folding the same ids in the same order
        assert _rel(two.global_models[ds].arena.cpu(), one.global_models[ds].arena.cpu()) <= 1e-3, ds
    a, b = (s.results[1]["clients_updated"]["loss"]["avg"] for s in (one, two))
    assert abs(a - b) <= 1e-4 * abs(a)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case,placement", [("fedcola", "reference"), ("fediot", "balanced"), ("fedprox", "reference")])
def test_two_rank_nccl_round_equals_one_gpu_lerp_round(case, placement, two_gpus, tmp_path):
    """The 2-rank round (client sharding, closed-form partial sums, NCCL all-reduce of the compact staging buffer,
    scatter, aux refresh) against the 1-GPU sequential-lerp aggregation of THE SAME trained client arenas: <= 1e-6
    norm-wise per tensor.  (Two separate training runs differ by more than that on their own — fp32 atomics reorder
    sums and a flipped bf16 rounding propagates — so the trained arenas are taken from the ranks, not re-trained;
    the independently re-trained 1-GPU round is compared at the run-to-run level, 1e-3.)"""
    from test_round_gpu import our_round
    out = str(tmp_path / "snap")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_nccl_round_worker.py"), case, out, placement]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:]
    snaps = [torch.load(f"{out}.rank{k}") for k in range(2)]
    assert snaps[0]["ids"] == snaps[1]["ids"] and snaps[0]["owner"] == snaps[1]["owner"]
    assert abs(snaps[0]["loss"] - snaps[1]["loss"]) == 0.0                     # every rank logs the same round
    assert sorted(set(snaps[0]["owner"].values())) == [0, 1]
    for ds in snaps[0]["new"]:
        assert torch.equal(snaps[0]["new"][ds], snaps[1]["new"][ds]), "ranks disagree after the all-reduce"
    assert snaps[0]["allreduce_bytes"] <= snaps[0]["arena_bytes"]              # only aggregated segments cross NVLink
    # ---- the same trained arenas through the 1-GPU bit-exact path
    d0 = two_gpus[0]
    one, ids, datasets, _ = our_round(case, d0, client_devices=["cuda:0"])     # an independent 1-GPU round
    assert list(ids) == snaps[0]["ids"]
    retrained = {ds: one.global_models[ds].arena.clone() for ds in datasets}
    for ds in datasets:
        one.global_models[ds].arena.copy_(snaps[0]["old"][ds].to(d0))
    trained = {**snaps[0]["clients"], **snaps[1]["clients"]}
    assert sorted(trained) == sorted(ids)
    for i in ids:
        c = one.clients[i]
        c.download(one.global_models)
        c.model.to(d0)
        c.model.arena.copy_(trained[i].to(d0))
    one._place(list(ids))
    one._aggregate_datasets(list(datasets), list(ids), snaps[0]["sizes"])
    if one.args.with_aux:
        one._refresh_aux()
    torch.cuda.synchronize()
    for ds in datasets:
        spec = one.global_models[ds].spec
        a, b = snaps[0]["new"][ds].numpy(), one.global_models[ds].arena.cpu().numpy()
        for s in spec.unique_segments():
            x, y = a[s.offset:s.offset + s.numel].astype(np.float64), b[s.offset:s.offset + s.numel].astype(np.float64)
            assert np.linalg.norm(x - y) <= 1e-6 * max(np.linalg.norm(y), 1e-30) + 1e-9, (ds, s.key)
        assert _rel(snaps[0]["new"][ds], retrained[ds].cpu()) <= 1e-3, ds
    ref_loss = one.results[1]["clients_updated"]["loss"]["avg"]
    assert abs(snaps[0]["loss"] - ref_loss) <= 1e-3 * abs(ref_loss)
