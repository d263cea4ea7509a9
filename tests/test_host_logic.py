"""CPU: host-side tables the native driver consumes — arena layout, work-item (chunk) tables, operand/aux tables and
the DropPath scale generator — checked for the invariants the kernels rely on."""
import numpy as np
import pytest
import torch

from fedcola_b200 import runtime
from fedcola_b200.arena import MatSpec, BLOCK_ROLES
from helpers import make_spec


def specs():
    out = []
    for ds in ("CIFAR100", "AG_NEWS", "Flickr30k"):
        for sp, sc, aux in (("none", "dataset", False), ("attn", "modality", True), ("attn", "all", False)):
            out.append(make_spec(ds, sp, sc, with_aux=aux, aux_trained=aux))
    return out


@pytest.mark.parametrize("spec", specs(), ids=lambda s: f"{s.modalities}-{s.share_scope}-aux{int(s.with_aux)}")
def test_arena_layout_invariants(spec):
    """Every tensor starts on a 128-byte boundary, unique segments do not overlap and fill `total` (up to padding),
    aliases (share_scope=all) point at an existing segment of the same shape."""
    uniq = spec.unique_segments()
    end = 0
    for s in uniq:
        assert s.offset % 32 == 0, s.key                     # 32 floats = 128 bytes
        assert s.offset >= end, s.key
        assert s.numel == int(np.prod(s.shape))
        end = s.offset + s.numel
    assert end <= spec.total and spec.total % 32 == 0
    by_key = {s.key: s for s in spec.segments}
    for s in spec.segments:
        if s.alias_of is not None:
            t = by_key[s.alias_of]
            assert (s.offset, s.numel, s.shape) == (t.offset, t.numel, t.shape) and t.alias_of is None
    assert spec.n_params() == sum(s.numel for s in uniq)
    # required_params()/aux_params() partitions (mome.py:844-878)
    req, aux = set(spec.required_keys()), set(spec.aux_keys())
    if spec.with_aux and None in spec.modalities:
        assert aux and not (req & aux)
        assert all("aux" in k for k in aux)
    assert all("cross_modal_scale" not in k for k in req) or not spec.with_aux


@pytest.mark.parametrize("spec", specs()[:6], ids=lambda s: f"{s.modalities}-{s.share_scope}-aux{int(s.with_aux)}")
def test_chunk_table_covers_trainable_elements_exactly_once(spec):
    plan = runtime.ModelPlan(spec)
    rg = {s.key: s.requires_grad for s in spec.segments}
    table, nseg = plan.chunk_table(rg)
    segs = [s for s in spec.unique_segments() if rg[s.key]]
    assert nseg == len(segs)
    cover = np.zeros(spec.total, dtype=np.int32)
    for off, ln, seg in zip(table["off"], table["len"], table["seg"]):
        assert 0 < ln <= plan.chunk and off % 4 == 0
        s = segs[int(seg)]
        assert s.offset <= off and off + ln <= s.offset + s.numel
        cover[off:off + ln] += 1
    want = np.zeros(spec.total, dtype=np.int32)
    for s in segs:
        want[s.offset:s.offset + s.numel] = 1
    assert np.array_equal(cover, want)
    assert plan.chunk_table(rg)[0] is table                                   # cached per signature
    frozen = dict(rg)
    frozen[segs[0].key] = False
    t2, n2 = plan.chunk_table(frozen)
    assert n2 == nseg - 1 and int(t2["len"].sum()) == int(table["len"].sum()) - segs[0].numel


@pytest.mark.parametrize("spec", specs()[:6], ids=lambda s: f"{s.modalities}-{s.share_scope}-aux{int(s.with_aux)}")
def test_operand_and_aux_tables(spec):
    """One bf16 operand slot per Linear (64-element aligned, disjoint), aux entries exactly for the Linears that
    carry (aux_weight, cross_modal_scale), block role offsets point into the arena."""
    plan = runtime.ModelPlan(spec)
    slots = sorted((int(r["dst_off"]), int(r["rows"]) * int(r["cols"])) for r in plan.prep_table)
    end = 0
    for dst, n in slots:
        assert dst % 64 == 0 and dst >= end
        end = dst + n
    assert end <= plan.operand_elems
    n_enc = sum(m is not None for m in spec.modalities)
    n_lin = 4 * spec.depth * n_enc + (1 if spec.modalities[0] is not None else 0)      # + patch projection
    assert len(plan.prep_table) == n_lin
    expect_aux = 4 * spec.depth * n_enc if spec.has_aux else 0
    assert len(plan.aux_table) == expect_aux
    for e in range(2):
        for j in range(spec.depth):
            for k, role in enumerate(BLOCK_ROLES):
                off = plan.desc.blk[e][j][k]
                if spec.modalities[e] is None:
                    assert off == -1
                elif off >= 0:
                    assert off < spec.total and off % 32 == 0, role


def test_droppath_scales_follow_the_stochastic_depth_rule():
    """timm DropPath as restated in oracle/ref_shim.py (parity unpinned: timm is not in the image): per block branch a
    Bernoulli(keep) mask per sample divided by keep, keep = 1 - linspace(0, rate, depth)[j]; the 'reference' mode
    consumes the generator exactly like the module sequence (encoder 0 first, attn branch then mlp branch)."""
    spec = MatSpec(embed_dim=64, depth=4, num_heads=1, modalities=("img", "txt"), num_classes=(None, None),
                   tasks=("rtv", "rtv"), vocab_size=64, max_text_len=8, drop_path_rate=0.3)
    B = 16
    assert runtime.droppath_scales(spec, B, "cpu", training=False) is None
    torch.manual_seed(5)
    got = runtime.droppath_scales(spec, B, "cpu", training=True, mode="reference")
    torch.manual_seed(5)
    dpr = [x.item() for x in torch.linspace(0, 0.3, 4)]
    for e in range(2):
        for j, r in enumerate(dpr):
            for br in range(2):
                if r <= 0:
                    assert torch.all(got[e, j, br] == 1)
                    continue
                want = torch.empty(B, 1, 1).bernoulli_(1 - r).div_(1 - r).view(B)
                assert torch.equal(got[e, j, br], want), (e, j, br)
    fused = runtime.droppath_scales(spec, B, "cpu", training=True, mode="fused")
    keep = torch.tensor([1 - r for r in dpr]).view(1, -1, 1, 1)
    vals = fused * keep
    assert torch.all((vals == 0) | ((vals - 1).abs() < 1e-6)) and torch.all(fused[:, 0] == 1)
    zero = MatSpec(embed_dim=64, depth=2, num_heads=1, drop_path_rate=0.0)
    assert runtime.droppath_scales(zero, B, "cpu", training=True) is None


def test_attention_bias_gradient_identities():
    """The identities the attention backward uses for the qkv bias gradient (csrc/attention.cu): with P = softmax(QK^T)
    row-stochastic, sum_k dV[k] = sum_q dO[q] (an all-ones row of P^T yields it), and sum_k dK[k] = 0 exactly
    (the scores are invariant to a shift of every key by the same vector)."""
    torch.manual_seed(0)
    B, H, N, D = 2, 3, 37, 64
    q, k, v = (torch.randn(B, H, N, D, dtype=torch.float64, requires_grad=True) for _ in range(3))
    do = torch.randn(B, H, N, D, dtype=torch.float64)
    o = torch.softmax((q * D ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v
    o.backward(do)
    assert torch.allclose(v.grad.sum(2), do.sum(2), rtol=1e-12, atol=1e-12)
    assert k.grad.sum(2).abs().max() < 1e-12 * k.grad.abs().max() * N
    # and the Q third is what the kernel reduces explicitly: nothing special about it
    assert q.grad.sum(2).abs().max() > 1e-3


def test_download_reuses_released_model_objects():
    """client.download() is `copy.deepcopy(models[dataset])` (fedavgclient.py:156); a model object released at the end of
    a round is refilled instead of rebuilt, and must be indistinguishable from a fresh deep copy: same values, same
    requires_grad flags, same training flag, no gradients, no shared storage with the global model."""
    import copy
    import types
    from fedcola_b200.client.fedavgclient import FedavgClient
    from fedcola_b200.models.mome import ModalityAgnosticTransformer
    torch.manual_seed(0)
    g = ModalityAgnosticTransformer(img_size=32, patch_size=16, embed_dim=32, depth=2, num_heads=2, modalities=["img", None],
                                    num_classes=[10, None], tasks=["cls", None], with_aux=True)
    c = FedavgClient.__new__(FedavgClient)
    c.dataset, c.device, c.model = "CIFAR100", "cpu", None
    models = {"CIFAR100": g}
    c.download(models)
    first = c.model
    assert first is not g and first.spec is g.spec
    assert first.arena.data_ptr() != g.arena.data_ptr() and torch.equal(first.arena, g.arena)
    # train the copy a bit, leave junk behind, then hand it back
    first.arena.add_(1.0)
    next(iter(first.parameters())).grad = torch.ones_like(next(iter(first.parameters())))
    for p in first.parameters():
        p.requires_grad_(False)
    first.eval()
    c.release_model(models)
    assert c.model is None and g._shell_pool == [first]
    # the global moves on; one block gets frozen on the server side
    g.arena.mul_(0.5)
    frozen = [k for k in g._params_by_key if k.startswith("blockses.0.1.")]
    for k in frozen:
        g._params_by_key[k].requires_grad_(False)
    g.train()
    c.download(models)
    assert c.model is first and g._shell_pool == []
    fresh = copy.deepcopy(g)
    assert torch.equal(c.model.arena, fresh.arena)
    assert c.model.training == fresh.training
    assert all(p.grad is None for p in c.model.parameters())
    for (k, p), (k2, q) in zip(c.model.named_parameters(), fresh.named_parameters()):
        assert k == k2 and p.requires_grad == q.requires_grad and torch.equal(p, q)
    assert not hasattr(fresh, "_shell_pool") or "_shell_pool" not in fresh.__dict__
    # a different architecture is never pooled under this global
    other = ModalityAgnosticTransformer(img_size=32, patch_size=16, embed_dim=32, depth=1, num_heads=2, modalities=["img", None],
                                        num_classes=[10, None], tasks=["cls", None])
    c2 = FedavgClient.__new__(FedavgClient)
    c2.dataset, c2.device, c2.model = "CIFAR100", "cpu", other
    c2.release_model(models)
    assert c2.model is None and g._shell_pool == []
