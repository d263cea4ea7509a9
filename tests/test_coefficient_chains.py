"""CPU: the known-answer mixing-coefficient chains of SURVEY.md 8(a') — generated there from the unmodified
reference — against the planner (`aggregation._client_coefs`) and the oracle (`fedcola_oracle.coefficients`).
Setup: datasets [CIFAR100, AG_NEWS, Flickr30k, Coco], three sampled clients 0: img n=16, 1: txt n=24,
2: img+txt n=16, out_modality_scales all 1.  A chain lists, for one (global, key), the clients that are folded in
(ascending id) with their coefficient; clients with coefficient 0 or without the key are skipped."""
from fractions import Fraction as F

import pytest

from fedcola_b200 import aggregation as agg
from oracle import fedcola_oracle as O
from helpers import make_spec

MODS = ["img", "txt", "img+txt", "img+txt"]
DS = {"CIFAR100": ("img", "cls"), "AG_NEWS": ("txt", "cls"), "Flickr30k": ("img+txt", "rtv")}
CLIENTS = [(0, "CIFAR100", "img", "cls", 16), (1, "AG_NEWS", "txt", "cls", 24), (2, "Flickr30k", "img+txt", "img+txt", 16)]

QKV0, QKV1 = "blockses.0.0.attn.qkv.weight", "blockses.1.0.attn.qkv.weight"
FC10, FC11 = "blockses.0.0.mlp.fc1.weight", "blockses.1.0.mlp.fc1.weight"
N10 = "blockses.0.0.norm1.weight"

# (shared_param, share_scope, compensation, with_aux, global dataset, key) -> [(client, coefficient)]
CHAINS = [
    ("attn", "modality", True, True, "CIFAR100", QKV0, [(0, F(16, 32)), (2, F(16, 32))]),
    ("attn", "modality", True, True, "CIFAR100", FC10, [(0, F(16, 32))]),                 # dataset scope, damped
    ("attn", "modality", True, True, "AG_NEWS", QKV1, [(1, F(24, 40)), (2, F(16, 40))]),
    ("attn", "modality", True, True, "Flickr30k", QKV0, [(0, F(16, 56)), (2, F(16, 56))]),
    ("attn", "modality", True, True, "Flickr30k", QKV1, [(1, F(24, 56)), (2, F(16, 56))]),
    ("attn", "modality", True, True, "Flickr30k", FC10, [(2, F(16, 56))]),                # damped
    ("attn", "modality", False, True, "CIFAR100", QKV0, [(0, F(1, 2)), (2, F(1, 2))]),
    ("attn", "modality", False, True, "CIFAR100", FC10, [(0, F(1))]),
    ("attn", "modality", False, True, "Flickr30k", QKV0, [(0, F(16, 56)), (2, F(16, 56))]),
    ("attn", "modality", False, True, "Flickr30k", FC10, [(2, F(1))]),
    ("blocks", "modality_exact", False, False, "CIFAR100", FC10, [(0, F(1, 2)), (2, F(1, 2))]),
    ("blocks", "modality_exact", False, False, "CIFAR100", N10, [(0, F(1, 2)), (2, F(1, 2))]),
    ("blocks", "modality_exact", False, False, "CIFAR100", QKV0, [(0, F(1))]),
    ("blocks", "modality_exact", False, False, "AG_NEWS", FC11, [(1, F(6, 10)), (2, F(4, 10))]),
    ("blocks", "modality_exact", False, False, "Flickr30k", FC10, [(0, F(1, 2)), (2, F(1, 2))]),
    ("blocks", "modality_exact", False, False, "Flickr30k", FC11, [(1, F(6, 10)), (2, F(4, 10))]),
    ("blocks", "modality_exact", False, False, "Flickr30k", QKV0, [(2, F(1))]),
    ("blocks", "modality_exact", False, False, "Flickr30k", "norm.weight", [(2, F(1))]),
    ("attn", "all", False, False, "CIFAR100", QKV0, [(0, F(16, 56)), (1, F(24, 56)), (2, F(16, 56))]),
    ("attn", "all", False, False, "AG_NEWS", QKV1, [(0, F(16, 56)), (1, F(24, 56)), (2, F(16, 56))]),
    ("attn", "all", False, False, "Flickr30k", QKV0, [(0, F(16, 56)), (1, F(24, 56)), (2, F(16, 56))]),
    ("attn", "all", False, False, "CIFAR100", FC10, [(0, F(1))]),
]


def _setup(sp, sc, aux):
    specs = {ds: make_spec(ds, sp, sc, with_aux=aux) for ds in DS}
    names = []
    for s in specs.values():
        for k in s.keys():
            if k not in names:
                names.append(k)
    scope = agg.init_param_scope(names, sp, sc)
    clients = [agg.ClientCtx(i, ds, m, t, n, specs[ds], None) for i, ds, m, t, n in CLIENTS]
    return specs, scope, clients


@pytest.mark.parametrize("sp,sc,comp,aux,gds,key,chain", CHAINS,
                         ids=[f"{c[0]}-{c[1]}-{'comp' if c[2] else 'nocomp'}-{c[4]}-{c[5]}" for c in CHAINS])
def test_known_answer_chain(sp, sc, comp, aux, gds, key, chain):
    specs, scope, clients = _setup(sp, sc, aux)
    gm, gt = DS[gds]
    g = agg.GlobalCtx(gds, gm, gt, 1, specs[gds], None, None)
    pm = agg.get_name_modality(key, MODS)
    planner = agg._client_coefs(scope[key], pm, g, clients, MODS, sc, comp, False)
    uploads = {c.id: set(agg.upload_keys(c.spec, aux, c.modality)) for c in clients}
    got = [(c.id, planner[c.id]) for c in clients if key in uploads[c.id] and planner[c.id] != 0]
    assert [i for i, _ in got] == [i for i, _ in chain]
    for (_, v), (_, want) in zip(got, chain):
        assert v == float(want), (v, want)          # the reference computes size / total in Python doubles
    # the oracle restatement gives the same numbers
    oc = O.coefficients([key], scope, {c.id: dict(dataset=c.dataset, modality=c.modality, task=c.task) for c in clients},
                        {c.id: c.size for c in clients}, gds, gm, gt, 1, MODS, sc, comp, False)
    for i, want in chain:
        assert oc[key][i] == float(want)
