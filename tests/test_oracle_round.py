"""CPU: the oracle's full-round restatement (oracle/round_oracle.py) against golden vectors of the UNMODIFIED
reference's `FedavgServer.update()` / `FedproxServer.update()` (tests/golden/train_golden.npz, `round/*`)."""
import os
import random

import numpy as np
import pytest
import torch

import helpers as H
from oracle.round_oracle import OracleServer

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz"))


def oracle_round(case):
    args, cds, datasets = H.round_args(case)
    specs = {ds: H.round_global_spec(case, ds) for ds in datasets}
    init = {ds: H.state_dict_of(specs[ds], H.fill_arena(specs[ds], 100 + i)) for i, ds in enumerate(datasets)}
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    server = OracleServer(args, cds, specs, init)
    server.round = 1
    ids = server.update()
    return server, ids, datasets, init


@pytest.mark.parametrize("case", sorted(H.ROUND_CASES))
def test_oracle_round_matches_reference(case):
    server, ids, datasets, init = oracle_round(case)
    assert list(ids) == list(GOLD[f"round/{case}/ids"])
    sizes = np.array([server.last_sizes[i] for i in ids], dtype=float)
    losses = np.array([server.last_losses[i] for i in ids], dtype=float)
    np.testing.assert_allclose(losses.dot(sizes) / sizes.sum(), GOLD[f"round/{case}/loss_avg"], rtol=1e-5)
    for ds in datasets:
        for k, v in server.globals[ds].params.items():
            np.testing.assert_allclose(H.subsample(v.numpy(), 7), GOLD[f"round/{case}/{ds}:{k}"], rtol=2e-4, atol=2e-6,
                                       err_msg=f"{ds}:{k}")
