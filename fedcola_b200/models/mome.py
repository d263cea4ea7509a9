"""Drop-in `ModalityAgnosticTransformer` over a flat fp32 arena (mirror of /root/reference/src/models/mome.py).

Same constructor arguments, `state_dict()` key names/order, `required_params()`, `aux_params()`,
`sync_shared_weights()`, `forward([img|None, ids|None], feat_out=False)` contract and model factories as
the reference (mome.py:671-1033) — but every parameter is a view into ONE contiguous fp32 buffer
(`fedcola_b200.arena.MatSpec`), and forward/backward run in the native sm_100a step driver
(csrc/mat_driver.cu) instead of ~16 eager kernels per block.

Initial weights are bit-identical to the reference for the same torch seed: the constructor draws from
torch's RNG by instantiating the same torch layers in the same order as the reference's `__init__`
(mome.py:713-769) and copies them into the arena.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn

from ..arena import MatSpec

_REGISTRY = {}


def register_model(fn):
    """timm.models.registry.register_model stand-in (mome.py:35,924)."""
    _REGISTRY[fn.__name__] = fn
    return fn


def create_model(model_str, pretrained=False, **kwargs):
    """timm.create_model stand-in used by FedavgServer._init_model (fedavgserver.py:151-155)."""
    if model_str not in _REGISTRY:
        raise RuntimeError(f"Unknown model ({model_str})")
    return _REGISTRY[model_str](pretrained=pretrained, **kwargs)


class _Node(nn.Module):
    """Parameter container; children named like the reference's sub-modules ('0', 'attn', 'qkv', ...)."""

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())


class ModalityAgnosticTransformer(nn.Module):
    def __init__(self, modalities, num_classes, tasks, shared_param="none", share_scope="dataset",
                 colearn_param="none", img_size=224, patch_size=16, in_chans=3, embed_dim=768, drop_rate=0.0,
                 num_heads=12, vocab_size=30522, max_text_len=40, mlp_ratio=4, qkv_bias=True, qk_scale=None,
                 attn_drop_rate=0.0, drop_path_rate=0.0, depth=12, shared_start_index=-1,
                 layer_scale_init_values=None, _init=True, _device=None, **kwargs):
        super().__init__()
        if not qkv_bias or qk_scale is not None or attn_drop_rate or drop_rate or layer_scale_init_values:
            raise NotImplementedError("fedcola_b200 supports the configuration every reference factory uses: "
                                      "qkv_bias=True, no qk_scale, no attention/projection dropout, no LayerScale")
        self.spec = MatSpec(embed_dim=embed_dim, depth=depth, num_heads=num_heads, modalities=tuple(modalities),
                            num_classes=tuple(num_classes), tasks=tuple(tasks), vocab_size=vocab_size,
                            max_text_len=max_text_len, img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                            mlp_ratio=mlp_ratio, drop_path_rate=float(drop_path_rate),
                            with_aux=kwargs.get("with_aux", False), aux_trained=kwargs.get("aux_trained", False),
                            aux_attn_only=kwargs.get("aux_attn_only", False),
                            aux_mlp_only=kwargs.get("aux_mlp_only", False),
                            share_scope="dataset",        # aliases appear only after sync_shared_weights()
                            shared_param=shared_param, colearn_param=colearn_param)
        self.embed_dim = embed_dim
        self.with_aux = self.spec.with_aux
        self.aux_trained = self.spec.aux_trained
        self.aux_attn_only = self.spec.aux_attn_only
        self.aux_mlp_only = self.spec.aux_mlp_only
        self.shared_start_index = depth if shared_start_index == -1 else shared_start_index
        self.shared_param = shared_param
        self.scope = share_scope
        self.colearn_param = colearn_param
        self.modalities = list(modalities)
        self.num_heads = num_heads
        self.precision = kwargs.get("precision", "bf16")
        self._arena = torch.zeros(self.spec.total, dtype=torch.float32, device=_device or "cpu")
        self._runtime = None            # native step driver state (lazily created on first CUDA forward)
        self._bind()
        if _init:
            self._reference_init()

    # ------------------------------------------------------------------------------------------
    # arena <-> nn.Parameter plumbing
    # ------------------------------------------------------------------------------------------
    def _bind(self):
        """(Re)create the module tree so that state_dict()/named_parameters() carry the reference's names."""
        old_flags = {k: p.requires_grad for k, p in self.named_parameters()} if len(self._modules) else {}
        for name in list(self._modules.keys()):
            del self._modules[name]
        made = {}
        for seg in self.spec.segments:
            parts = seg.key.split(".")
            if seg.alias_of is not None:
                continue          # bound below, once the main encoder exists
            node = self
            for p in parts[:-1]:
                if p not in node._modules or node._modules[p] is None:
                    node._modules[p] = _Node()
                node = node._modules[p]
            view = self._arena[seg.offset:seg.offset + seg.numel].view(seg.shape)
            param = nn.Parameter(view, requires_grad=old_flags.get(seg.key, seg.requires_grad))
            node._parameters[parts[-1]] = param
            made[seg.key] = param
        for seg in self.spec.segments:
            if seg.alias_of is not None:
                # share_scope == 'all': blockses.<none_idx> *is* blockses.<main_idx> (mome.py:824-827)
                dst, src = seg.key.split(".")[1], seg.alias_of.split(".")[1]
                self._modules["blockses"]._modules[dst] = self._modules["blockses"]._modules[src]
        # None placeholders, as in the reference's ModuleLists
        for grp in ("embeddings", "blockses", "heads"):
            if grp not in self._modules:
                self._modules[grp] = _Node()
            for i in range(2):
                if str(i) not in self._modules[grp]._modules:
                    self._modules[grp]._modules[str(i)] = None
            self._modules[grp]._modules = dict(sorted(self._modules[grp]._modules.items()))
        # registration order must follow the reference: embeddings, blockses, norm, heads
        self._modules = {k: self._modules[k] for k in ("embeddings", "blockses", "norm", "heads")}
        self._params_by_key = made
        self.__dict__["_train_synced"] = False

    def _apply(self, fn, recurse=True):
        new = fn(self._arena)
        if new is not self._arena:
            self._arena = new
            self._runtime = None
            for seg in self.spec.unique_segments():
                p = self._params_by_key[seg.key]
                p.data = self._arena[seg.offset:seg.offset + seg.numel].view(seg.shape)
                p.grad = None
        return self

    def __deepcopy__(self, memo):
        new = ModalityAgnosticTransformer.__new__(ModalityAgnosticTransformer)
        nn.Module.__init__(new)
        for k, v in self.__dict__.items():
            if k in ("_parameters", "_buffers", "_modules", "_arena", "_runtime", "_params_by_key", "_shell_pool", "_train_synced") or \
                    k.startswith("_forward") or k.startswith("_backward") or k.startswith("_state_dict") or \
                    k.startswith("_load_state_dict"):
                continue
            # the spec is immutable once the model is built (sync_shared_weights runs in __init__), so copies share it
            new.__dict__[k] = v if k == "spec" else copy.deepcopy(v, memo)
        new._arena = self._arena.clone()
        new._runtime = None
        new._bind()
        for k, p in self._params_by_key.items():
            new._params_by_key[k].requires_grad_(p.requires_grad)
        new.training = self.training
        return new

    def train(self, mode=True):
        """nn.Module.train, without walking the ~300 placeholder sub-modules when nothing changes (a round calls this
        on every client model; the recursion costs ~2 ms of interpreter time per call)."""
        if self.training == bool(mode) and getattr(self, "_train_synced", False):
            return self
        super().train(mode)
        self.__dict__["_train_synced"] = True
        return self

    def refill_from(self, other):
        """Make this model a copy of `other` (same spec) without rebuilding the module tree: one arena copy (device to
        device, also across GPUs) plus the requires_grad flags and the training flag.  What `copy.deepcopy(other)`
        yields, for a model object that is being reused (client/fedavgclient.py::download)."""
        if other.spec is not self.spec and other.spec.signature != self.spec.signature:
            raise ValueError("refill_from: the two models differ in architecture")
        self._arena.copy_(other._arena, non_blocking=True)
        mine = self._params_by_key
        for k, p in other._params_by_key.items():
            q = mine[k]
            if q.requires_grad != p.requires_grad:
                q.requires_grad_(p.requires_grad)
            q.grad = None
        self._runtime = None
        self.train(other.training)
        return self

    @property
    def arena(self):
        """The flat fp32 parameter buffer (a torch tensor on the model's device)."""
        return self._arena

    @property
    def device(self):
        return self._arena.device

    # ------------------------------------------------------------------------------------------
    # initialisation: same torch RNG call sequence as the reference constructor
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _reference_init(self):
        sp, d = self.spec, self.spec.embed_dim
        P = self._params_by_key
        for i, m in enumerate(sp.modalities):
            if m == "img":       # PatchEmbed conv (mome.py:252-258); pos_embed / cls_token stay zero (:592-595)
                conv = nn.Conv2d(sp.in_chans, d, kernel_size=sp.patch_size, stride=sp.patch_size, bias=True)
                P[f"embeddings.{i}.embed.proj.weight"].copy_(conv.weight)
                P[f"embeddings.{i}.embed.proj.bias"].copy_(conv.bias)
            elif m == "txt":     # HF BertEmbeddings (mome.py:618-626)
                from transformers.models.bert.modeling_bert import BertConfig, BertEmbeddings
                be = BertEmbeddings(BertConfig(vocab_size=sp.vocab_size, hidden_size=d,
                                               max_position_embeddings=sp.max_text_len, hidden_dropout_prob=0.0,
                                               position_embedding_type="absolute"))
                t = f"embeddings.{i}.text_embeddings."
                P[t + "word_embeddings.weight"].copy_(be.word_embeddings.weight)
                P[t + "position_embeddings.weight"].copy_(be.position_embeddings.weight)
                P[t + "token_type_embeddings.weight"].copy_(be.token_type_embeddings.weight)
                P[t + "LayerNorm.weight"].copy_(be.LayerNorm.weight)
                P[t + "LayerNorm.bias"].copy_(be.LayerNorm.bias)
        hid = d * sp.mlp_ratio
        for i, m in enumerate(sp.modalities):
            if m is None:
                continue
            for j in range(sp.depth):
                p = f"blockses.{i}.{j}."
                P[p + "norm1.weight"].fill_(1.0)
                P[p + "norm2.weight"].fill_(1.0)
                for lname, shp in (("attn.qkv", (3 * d, d)), ("attn.proj", (d, d)), ("mlp.fc1", (hid, d)),
                                   ("mlp.fc2", (d, hid))):
                    lin = nn.Linear(shp[1], shp[0])
                    P[p + lname + ".weight"].copy_(lin.weight)
                    P[p + lname + ".bias"].copy_(lin.bias)
        P["norm.weight"].fill_(1.0)
        for i, t in enumerate(sp.tasks):
            if f"heads.{i}.head.weight" in P:
                lin = nn.Linear(d, sp.num_classes[i])
                P[f"heads.{i}.head.weight"].copy_(lin.weight)
                P[f"heads.{i}.head.bias"].copy_(lin.bias)
        if sp.has_aux:           # build_aux (mome.py:771-786): A = the original weight, W = a copy, s = 0;
            i = sp.main_idx      # CrossModalReparamLinear.__init__ draws (and discards) a fresh nn.Linear init
            for j in range(sp.depth):
                p = f"blockses.{i}.{j}."
                for lname in sp.aux_layer_names():
                    w = P[p + lname + ".weight"]
                    nn.Linear(w.shape[1], w.shape[0])
                    P[p + lname + ".aux_weight"].copy_(w)
                    P[p + lname + ".cross_modal_scale"].zero_()

    # ------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------
    def sync_shared_weights(self):
        """mome.py:818-842.  share_scope=='all' aliases the None encoder to the main one."""
        if self.scope == "all" and None in self.modalities and self.spec.share_scope != "all":
            flags = {k: p.requires_grad for k, p in self._params_by_key.items()}
            self.spec.share_scope = "all"
            self.spec._build()
            self._bind()
            for k, f in flags.items():
                self._params_by_key[k].requires_grad_(f)
        if self.colearn_param == "attn" and None not in self.modalities:
            raise NotImplementedError("colearn_param='attn' is not supported by the B200 path yet")
        # colearn_param == 'blocks' is a no-op in the reference (rebinding a loop variable, mome.py:833-836)

    def pretrain_vit(self, model_strs):
        raise NotImplementedError("pretrained timm checkpoints need network access (mome.py:788-816); "
                                  "load a state_dict with load_state_dict(..., strict=False) instead")

    def required_params(self):
        sd = self.state_dict()
        return {k: sd[k] for k in self.spec.required_keys()}

    def aux_params(self):
        if not self.with_aux:
            raise ValueError("No aux params.")
        sd = self.state_dict()
        return {k: sd[k] for k in self.spec.aux_keys()}

    def forward(self, x, feat_out=False):
        from .. import runtime
        return runtime.model_forward(self, x, feat_out)


def _factory(embed_dim, depth, num_heads):
    def make(pretrained, args, **kwargs):
        model = ModalityAgnosticTransformer(
            img_size=224, patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
            vocab_size=args.vocab_size, max_text_len=args.seq_len, drop_path_rate=args.dropout,
            shared_param=args.shared_param, share_scope=args.share_scope, colearn_param=args.colearn_param,
            precision=getattr(args, "precision", "bf16"), **kwargs)
        model.sync_shared_weights()
        if pretrained:
            model.pretrain_vit([None, None])
        return model
    return make


# the reference's factories (mome.py:924-1033) + the sizes BASELINE.json names that have no factory upstream
for _name, _cfg in {"mome_small_patch16": (384, 12, 6), "mome_tiny_patch16": (192, 12, 3),
                    "mome_small_patch16_224_in21k": (384, 12, 6), "mome_base_patch16_224_ours": (768, 12, 12),
                    "mome_toy_patch16_224": (4, 1, 2),        # mome.py:1016 (head_dim 2: constructs, cannot run here)
                    "mome_base_patch16": (768, 12, 12), "mome_d192_l4": (192, 4, 3),
                    "mome_d64_l2": (64, 2, 1)}.items():
    _f = _factory(*_cfg)
    _f.__name__ = _name
    register_model(_f)
