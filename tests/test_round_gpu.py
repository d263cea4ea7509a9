"""GPU: one full federated round through the drop-in servers/clients (fedcola_b200.server / .client) against
golden vectors of the UNMODIFIED reference's `server.update()` (tests/golden/train_golden.npz `round/*`).
Client sampling is bit-exact; losses and the global-model updates are within the bf16 budget (2e-2, L2 norm
of the update — see test_model_gpu.py for why norms)."""
import os
import random

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz"))
TOL = 2e-2


def our_round(case, cuda, **extra):
    from fedcola_b200.server import fedavgserver as fs
    from fedcola_b200.server.fedproxserver import FedproxServer
    from oracle.ref_shim import NullWriter
    fs.VOCAB_SIZES.update(H.TINY_VOCAB)
    args, cds, datasets = H.round_args(case)
    args.server_device = str(cuda)
    for k, v in extra.items():
        setattr(args, k, v)
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    S = FedproxServer if args.algorithm == "fedprox" else fs.FedavgServer
    server = S(args=args, writer=NullWriter(), server_dataset=(None, {}), client_datasets=cds,
               model_str=args.model_name)
    init = {}
    for i, ds in enumerate(datasets):
        spec = H.round_global_spec(case, ds)
        g = server.global_models[ds]
        assert g.spec.keys() == spec.keys()
        a = H.fill_arena(spec, 100 + i)
        g.arena.copy_(torch.from_numpy(a))
        init[ds] = H.state_dict_of(spec, a)
    server.round = 1
    ids = server.update()
    return server, ids, datasets, init


@pytest.mark.parametrize("case", sorted(H.ROUND_CASES))
def test_round_matches_reference(case, cuda):
    server, ids, datasets, init = our_round(case, cuda)
    assert list(ids) == list(GOLD[f"round/{case}/ids"])
    got_loss = server.results[1]["clients_updated"]["loss"]["avg"]
    ref_loss = float(GOLD[f"round/{case}/loss_avg"])
    assert abs(got_loss - ref_loss) <= TOL * abs(ref_loss), (got_loss, ref_loss)
    num = den = 0.0
    for ds in datasets:
        sd = {k: v.detach().cpu().numpy() for k, v in server.global_models[ds].state_dict().items()}
        for k, v in sd.items():
            ref = GOLD[f"round/{case}/{ds}:{k}"]
            base = H.subsample(init[ds][k], 7)
            d_ref, d_got = ref - base, H.subsample(v, 7) - base
            if "aux_weight" in k:      # refreshed from the other modality's global: a copy, compare values
                assert np.linalg.norm(H.subsample(v, 7) - ref) <= TOL * np.linalg.norm(ref) + 1e-6, (ds, k)
                continue
            num += float(np.sum((d_got - d_ref) ** 2))
            den += float(np.sum(d_ref ** 2))
    # SGD updates are linear in the gradients -> bf16 budget; Adam divides by |g| -> sign noise where g ~ 0
    adam = H.ROUND_CASES[case][0].get("optimizer", "SGD") == "AdamW"
    assert (num / max(den, 1e-30)) ** 0.5 <= (0.35 if adam else 2 * TOL)
    assert all(c.model is None for c in server.clients)            # _empty_client_models
    assert server.last_aggregation["bytes"] > 0


def test_round_threads_and_device_resident_data(cuda):
    """num_thread > 1 (per-client CUDA streams) and HBM-resident client data give the same round."""
    a, ids_a, datasets, _ = our_round("fedcola", cuda)
    b, ids_b, _, _ = our_round("fedcola", cuda, num_thread=3, data_resident="device")
    assert ids_a == ids_b
    for ds in datasets:
        x, y = a.global_models[ds].arena, b.global_models[ds].arena
        assert (x - y).norm() <= 1e-3 * x.norm()     # atomics reorder fp32 sums; nothing else differs


def test_upload_matches_aggregated_view(cuda):
    """client.upload() (API-compat path) equals what the fused aggregation reads: W + A*s, aux keys dropped."""
    from fedcola_b200.server import fedavgserver as fs
    server, ids, datasets, _ = our_round("fedcola", cuda)
    c = server.clients[0]
    c.download(server.global_models)
    sd = c.model.state_dict()
    k = "blockses.0.0.attn.qkv.weight"
    sd[k.replace("weight", "cross_modal_scale")].fill_(0.5)
    up = c.upload()
    assert not any("aux" in n or "cross_modal_scale" in n for n in up)
    assert torch.equal(up[k], sd[k] + sd[k.replace("weight", "aux_weight")] * 0.5)
