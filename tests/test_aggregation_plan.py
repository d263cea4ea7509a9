"""CPU: the host planner (coefficients, upload keys, job tables) against the oracle restatement of
FedavgServer._aggregate, bit-exact, for every share-scope mode.  The tables are executed by a numpy
interpreter (tests/helpers.run_plan_numpy) that follows csrc/aggregate.cu instruction by instruction."""
import numpy as np
import pytest
import torch

from fedcola_b200 import aggregation as agg
from oracle import fedcola_oracle as O
from helpers import AGG_CASES, build_agg_case, state_dict_of


def oracle_aggregate(gl, cl, scope, flags, fedavg=False):
    """New global state_dicts per dataset via the oracle (coefficients -> upload_merge -> sequential lerp)."""
    clients = {c.id: dict(dataset=c.dataset, modality=c.modality, task=c.task) for c in cl}
    sizes = {c.id: c.size for c in cl}
    ids = sorted(sizes)
    uploads = {c.id: O.upload_merge(state_dict_of(c.spec, c.arena.numpy()), flags["with_aux"], c.modality)
               for c in cl}
    out = {}
    for g in gl:
        names = g.spec.required_keys()
        coefs = O.coefficients(names, scope, clients, sizes, g.dataset, g.modality, g.task, g.out_modality_scale,
                               flags["args_modalities"], flags["share_scope_flag"], flags["compensation"], fedavg)
        sd = state_dict_of(g.spec, g.arena_in.numpy())
        final = {k: sd[k].copy() for k in names}
        out[g.dataset] = O.aggregate_lerp(final, uploads, coefs, ids)
    return out


@pytest.mark.parametrize("case", sorted(AGG_CASES))
def test_plan_matches_oracle_bit_exact(case):
    gl, cl, scope, flags = build_agg_case(case)
    expect = oracle_aggregate(gl, cl, scope, flags)
    plan = agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags)
    from helpers import run_plan_numpy
    run_plan_numpy(plan)
    for g in gl:
        got = state_dict_of(g.spec, g.arena_out.numpy())
        for k, v in expect[g.dataset].items():
            assert np.array_equal(got[k], v), (case, g.dataset, k)
        # keys outside required_params (aux_weight, cross_modal_scale) are untouched
        for s in g.spec.segments:
            if s.key not in expect[g.dataset] and s.alias_of is None:
                assert np.array_equal(got[s.key], state_dict_of(g.spec, g.arena_in.numpy())[s.key])


@pytest.mark.parametrize("case", ["fedcola_attn_modality_comp_aux", "attn_all", "attn_modality_scaled"])
def test_closed_form_matches_sequential(case):
    """WSUM mode (multi-GPU closed form) agrees with the sequential lerp to 1e-6 relative."""
    gl, cl, scope, flags = build_agg_case(case)
    expect = oracle_aggregate(gl, cl, scope, flags)
    plan = agg.AggregationPlan(gl, cl, scope, mode=agg.WSUM, **flags)
    from helpers import run_plan_numpy
    run_plan_numpy(plan)
    for g in gl:
        got = state_dict_of(g.spec, g.arena_out.numpy())
        for k, v in expect[g.dataset].items():
            np.testing.assert_allclose(got[k], v, rtol=1e-6, atol=1e-7, err_msg=f"{case} {g.dataset} {k}")


def test_closed_form_sharded_partial_sums():
    """Two ranks, each holding a shard of the clients: partial WSUMs add up to the single-rank result."""
    case = "attn_modality_scaled"
    gl, cl, scope, flags = build_agg_case(case)
    ref_gl, ref_cl, _, _ = build_agg_case(case)
    from helpers import run_plan_numpy
    run_plan_numpy(agg.AggregationPlan(ref_gl, ref_cl, scope, mode=agg.WSUM, **flags))
    total = [torch.zeros_like(g.arena_out) for g in gl]
    for rank in range(2):
        gl_r, cl_r, _, _ = build_agg_case(case)
        for g in gl_r:
            g.arena_out.zero_()
        for i, c in enumerate(cl_r):
            if i % 2 != rank:
                c.arena = None
        run_plan_numpy(agg.AggregationPlan(gl_r, cl_r, scope, mode=agg.WSUM, include_global_term=(rank == 0), **flags))
        for t, g in zip(total, gl_r):
            t += g.arena_out
    for t, g in zip(total, ref_gl):
        for k in g.spec.required_keys():
            s = g.spec.seg(k)
            np.testing.assert_allclose(t[s.offset:s.offset + s.numel].numpy(),
                                       g.arena_out[s.offset:s.offset + s.numel].numpy(), rtol=2e-6, atol=1e-7)


def test_known_answer_coefficient_chains():
    """SURVEY §8a' chains generated from the reference: (client, c) per (global, key)."""
    gl, cl, scope, flags = build_agg_case("fedcola_attn_modality_comp_aux")
    def chain(gi, key):
        g = gl[gi]
        pm = agg.get_name_modality(key, flags["args_modalities"])
        c = agg._client_coefs(scope[key], pm, g, cl, flags["args_modalities"], flags["share_scope_flag"],
                              flags["compensation"], False)
        return c
    assert chain(0, "blockses.0.0.attn.qkv.weight") == {0: 16 / 32, 1: 0.0, 2: 16 / 32}
    assert chain(0, "blockses.0.0.mlp.fc1.weight") == {0: 16 / 32, 1: 0.0, 2: 0.0}          # damped (F10b)
    assert chain(1, "blockses.1.0.attn.qkv.weight") == {0: 0.0, 1: 24 / 40, 2: 16 / 40}
    assert chain(2, "blockses.0.0.attn.qkv.weight") == {0: 16 / 56, 1: 24 / 56, 2: 16 / 56}  # txt counted, skipped later
    assert chain(2, "blockses.0.0.mlp.fc1.weight") == {0: 0.0, 1: 0.0, 2: 16 / 56}
    gl, cl, scope, flags = build_agg_case("fediot_blocks_modality_exact")
    def chain2(gi, key):
        pm = agg.get_name_modality(key, flags["args_modalities"])
        return agg._client_coefs(scope[key], pm, gl[gi], cl, flags["args_modalities"], flags["share_scope_flag"],
                                 flags["compensation"], False)
    assert chain2(0, "blockses.0.0.mlp.fc1.weight") == {0: 0.5, 1: 0.0, 2: 0.5}
    assert chain2(0, "blockses.0.0.attn.qkv.weight") == {0: 1.0, 1: 0.0, 2: 0.0}
    assert chain2(1, "blockses.1.0.mlp.fc1.weight") == {0: 0.0, 1: 0.6, 2: 0.4}
    assert chain2(2, "blockses.1.0.mlp.fc1.weight") == {0: 0.0, 1: 0.6, 2: 0.4}
    assert chain2(2, "norm.weight") == {0: 0.0, 1: 0.0, 2: 1.0}


def test_param_scope_rules():
    names = ["embeddings.0.pos_embed", "blockses.0.0.attn.qkv.weight", "blockses.0.0.mlp.fc1.weight",
             "blockses.0.0.norm1.weight", "norm.weight", "heads.0.head.weight"]
    s = agg.init_param_scope(names, "attn", "modality")
    assert [s[n] for n in names] == ["dataset", "modality", "dataset", "dataset", "dataset", "dataset"]
    s = agg.init_param_scope(names, "blocks", "modality_exact")
    assert [s[n] for n in names] == ["dataset", "dataset", "modality_exact", "modality_exact", "dataset", "dataset"]
    s = agg.init_param_scope(names, "mlp", "all")          # the 'mlp' branch is unreachable upstream
    assert set(s.values()) == {"dataset"}
    assert agg.init_param_scope(names, "blocks", "all") == O.init_param_scope(names, "blocks", "all")


def test_lerp_refuses_remote_clients():
    gl, cl, scope, flags = build_agg_case("attn_all")
    cl[1].arena = None
    with pytest.raises(ValueError):
        agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags)


# ---- golden vectors generated from the UNMODIFIED reference (oracle/make_golden.py) -------------------
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agg_hashes.json")


@pytest.mark.parametrize("case", sorted(AGG_CASES))
def test_oracle_and_plan_match_reference_golden(case):
    from helpers import run_plan_numpy, sha
    with open(GOLDEN) as f:
        golden = json.load(f)[case]
    gl, cl, scope, flags = build_agg_case(case)
    expect = oracle_aggregate(gl, cl, scope, flags)
    plan = agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags)
    run_plan_numpy(plan)
    for g in gl:
        got = state_dict_of(g.spec, g.arena_out.numpy())
        assert sorted(got) == sorted(golden[g.dataset])
        for k, h in golden[g.dataset].items():
            assert sha(got[k]) == h, ("plan", case, g.dataset, k)
            if k in expect[g.dataset]:
                assert sha(expect[g.dataset][k]) == h, ("oracle", case, g.dataset, k)


def test_memoised_plan_is_rebound_to_new_arenas():
    """The symbolic tables are cached on the structural signature of the round; a hit must address THIS round's
    arenas (fresh allocations, fresh values, deep-copied specs) and a different size pattern must miss."""
    import copy
    from helpers import run_plan_numpy
    case = "fedcola_attn_modality_comp_aux"
    gl, cl, scope, flags = build_agg_case(case)
    agg._PLAN_CACHE.clear()
    agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags)
    assert len(agg._PLAN_CACHE) == 1
    rng = np.random.default_rng(7)
    gl2 = [agg.GlobalCtx(g.dataset, g.modality, g.task, g.out_modality_scale, copy.deepcopy(g.spec),
                         torch.from_numpy(rng.standard_normal(g.spec.total).astype(np.float32)), torch.empty(g.spec.total))
           for g in gl]
    for g in gl2:
        g.arena_out.copy_(g.arena_in)
    cl2 = [agg.ClientCtx(c.id + 100, c.dataset, c.modality, c.task, c.size, copy.deepcopy(c.spec),
                         torch.from_numpy(rng.standard_normal(c.spec.total).astype(np.float32))) for c in cl]
    plan = agg.AggregationPlan(gl2, cl2, scope, mode=agg.LERP, **flags)
    assert len(agg._PLAN_CACHE) == 1, "same structure: cache hit"
    run_plan_numpy(plan)
    expect = oracle_aggregate(gl2, cl2, scope, flags)
    for g in gl2:
        got = state_dict_of(g.spec, g.arena_out.numpy())
        for k, v in expect[g.dataset].items():
            assert np.array_equal(got[k], v), (g.dataset, k)
    cl3 = [agg.ClientCtx(c.id, c.dataset, c.modality, c.task, c.size + (1 if i == 0 else 0), c.spec, c.arena)
           for i, c in enumerate(cl2)]
    for g in gl2:
        g.arena_out.copy_(g.arena_in)
    plan3 = agg.AggregationPlan(gl2, cl3, scope, mode=agg.LERP, **flags)
    assert len(agg._PLAN_CACHE) == 2, "different client sizes: new coefficients, new entry"
    run_plan_numpy(plan3)
    expect3 = oracle_aggregate(gl2, cl3, scope, flags)
    for g in gl2:
        got = state_dict_of(g.spec, g.arena_out.numpy())
        for k, v in expect3[g.dataset].items():
            assert np.array_equal(got[k], v), (g.dataset, k)


@pytest.mark.parametrize("seed", range(40))
def test_random_configurations_match_oracle_bit_exact(seed):
    """Seeded random scope / compensation / aux / scaling / client mixes: planner tables (executed like the kernel)
    equal the oracle's sequential lerp bit for bit; the closed form agrees to 1e-6."""
    from helpers import AGG_CASES, random_agg_case, run_plan_numpy
    name = f"random_{seed}"
    AGG_CASES[name] = random_agg_case(seed)
    try:
        gl, cl, scope, flags = build_agg_case(name)
        expect = oracle_aggregate(gl, cl, scope, flags)
        run_plan_numpy(agg.AggregationPlan(gl, cl, scope, mode=agg.LERP, **flags))
        for g in gl:
            got = state_dict_of(g.spec, g.arena_out.numpy())
            for k, v in expect[g.dataset].items():
                assert np.array_equal(got[k], v), (AGG_CASES[name], g.dataset, k)
        for g in gl:
            g.arena_out.copy_(g.arena_in)
        run_plan_numpy(agg.AggregationPlan(gl, cl, scope, mode=agg.WSUM, **flags))
        for g in gl:
            got = state_dict_of(g.spec, g.arena_out.numpy())
            for k, v in expect[g.dataset].items():
                np.testing.assert_allclose(got[k], v, rtol=1e-6, atol=1e-7, err_msg=f"{AGG_CASES[name]} {g.dataset} {k}")
    finally:
        del AGG_CASES[name]
