"""CPU, world_size 2, gloo: the multi-rank aggregation path (client sharding rule + closed-form partial sums
+ one all-reduce), with the plan tables executed by the numpy interpreter instead of the CUDA kernel.
Every rank must end with the same global arenas, equal (<= 1e-6 relative) to the single-process sequential
lerp of the reference."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fedcola_b200 import aggregation as agg
    from helpers import build_agg_case, run_plan_numpy
    gl, cl, scope, flags = build_agg_case(case)
    for pos, c in enumerate(sorted(cl, key=lambda c: c.id)):
        if agg.shard_owner(pos, world) != rank:
            c.arena = None                       # trained on the other rank: only its meta data is known here
    agg.sharded_aggregate(gl, cl, scope, flags, dist, rank, execute=run_plan_numpy)
    torch.save([g.arena_in for g in gl], os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["fedcola_attn_modality_comp_aux", "attn_modality_scaled", "attn_all"])
def test_two_rank_aggregation_matches_sequential(case, tmp_path):
    from test_aggregation_plan import oracle_aggregate
    from helpers import build_agg_case, state_dict_of
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    gl, cl, scope, flags = build_agg_case(case)
    expect = oracle_aggregate(gl, cl, scope, flags)
    for g, a0, a1 in zip(gl, r0, r1):
        assert torch.equal(a0, a1), "ranks disagree after the all-reduce"
        got = state_dict_of(g.spec, a0.numpy())
        for k, v in expect[g.dataset].items():
            np.testing.assert_allclose(got[k], v, rtol=2e-6, atol=1e-7, err_msg=f"{case} {g.dataset} {k}")
        old = state_dict_of(g.spec, g.arena_in.numpy())
        for s in g.spec.segments:                       # aux / scale keys are not aggregated
            if s.key not in expect[g.dataset] and s.alias_of is None:
                np.testing.assert_array_equal(got[s.key], old[s.key])


def test_shard_owner_rule():
    from fedcola_b200 import aggregation as agg
    assert [agg.shard_owner(i, 4) for i in range(9)] == [0, 1, 2, 3, 0, 1, 2, 3, 0]
    assert [agg.shard_owner(i, 1) for i in range(3)] == [0, 0, 0]
