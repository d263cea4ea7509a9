"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (/root/reference) in this container.

Used by oracle/make_golden.py (fixture generation) and by the CPU tests that pin the oracle
restatement against the real reference. `/root/reference` does not exist on the GPU box, so nothing
in `-m gpu` tests, `smoke()` or `bench.py` may import this module.

The reference imports six third-party modules that are not installed here (timm 0.9.12,
torchmultimodal, torchtext, medmnist, pycocotools, ml_collections).  Only four behaviours of
them are live on the hot path; they are stubbed below.

  * timm DropPath / to_2tuple / register_model / create_model     -> src/models/mome.py:29-37,213,223
  * torchmultimodal ContrastiveLossWithTemperature                -> src/criterions/__init__.py:3,8
    ("parity unpinned": restated from the published formula; identical to transformers' clip_loss)
"""
import collections.abc
import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("FEDCOLA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


class DropPath(nn.Module):
    """timm 0.9.12 stochastic depth, per-sample (restated; source not available offline)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return x * r


class ContrastiveLossWithTemperature(nn.Module):
    """torchmultimodal (unpinned) single-process path, restated."""

    def __init__(self, logit_scale=math.log(1 / 0.07), logit_scale_min=math.log(1), logit_scale_max=math.log(100)):
        super().__init__()
        self.lo, self.hi = logit_scale_min, logit_scale_max
        self.logit_scale = nn.Parameter(logit_scale * torch.ones([]))

    def forward(self, a, b):
        self.logit_scale.data.clamp_(self.lo, self.hi)
        t = torch.exp(self.logit_scale)
        lab = torch.arange(a.size(0), device=a.device)
        return (F.cross_entropy(a @ b.t() * t, lab) + F.cross_entropy(b @ a.t() * t, lab)) / 2


_REG = {}


def register_model(fn):
    _REG[fn.__name__] = fn
    return fn


def create_model(name, pretrained=False, **kw):
    return _REG[name](pretrained=pretrained, **kw)


_installed = False


def install():
    """Install the stubs and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import transformers  # noqa: F401  (must be imported BEFORE the timm stub exists)
    from transformers.models.bert.modeling_bert import BertConfig, BertEmbeddings  # noqa: F401

    _D = type("_D", (nn.Module,), {})
    _f = lambda *a, **k: None  # noqa: E731
    layers = _stub("timm.layers", PatchEmbed=_D, Mlp=_D, DropPath=DropPath, AttentionPoolLatent=_D, RmsNorm=_D,
                   PatchDropout=_D, SwiGLUPacked=_D, trunc_normal_=nn.init.trunc_normal_, lecun_normal_=_f,
                   resample_patch_embed=_f, resample_abs_pos_embed=_f, use_fused_attn=_f, get_act_layer=_f,
                   get_norm_layer=_f, LayerType=object)
    ml = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple, trunc_normal_=nn.init.trunc_normal_)
    mr = _stub("timm.models.registry", register_model=register_model)
    mm = _stub("timm.models", create_model=create_model, layers=ml, registry=mr)
    _stub("timm", layers=layers, models=mm, create_model=create_model)

    _stub("torchtext")
    sys.modules["torchtext"].datasets = _stub("torchtext.datasets")
    _stub("medmnist", INFO={})
    _stub("pycocotools")
    _stub("pycocotools.coco", COCO=object)
    _stub("ml_collections")
    try:
        import wandb  # noqa: F401
    except Exception:
        _stub("wandb")
    _stub("torchmultimodal")
    _stub("torchmultimodal.modules")
    _stub("torchmultimodal.modules.losses")
    _stub("torchmultimodal.modules.losses.contrastive_loss_with_temperature",
          ContrastiveLossWithTemperature=ContrastiveLossWithTemperature)

    sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import src  # noqa: F401
        import src.criterions  # noqa: F401  (patches torch.nn.ContrastiveLoss)
        import src.models.mome as mome

    @register_model
    def mome_d192_l4(pretrained, args, **kwargs):
        """BASELINE.json configs[0]: 4-layer d=192 model (no factory upstream)."""
        model = mome.ModalityAgnosticTransformer(
            img_size=224, patch_size=16, embed_dim=192, depth=4, num_heads=3, vocab_size=args.vocab_size,
            max_text_len=args.seq_len, drop_path_rate=args.dropout, shared_param=args.shared_param,
            share_scope=args.share_scope, colearn_param=args.colearn_param, **kwargs)
        model.sync_shared_weights()
        return model

    @register_model
    def mome_d64_l2(pretrained, args, **kwargs):
        """Tiny fixture model (2 layers, d=64, 1 head) for fast golden vectors."""
        model = mome.ModalityAgnosticTransformer(
            img_size=224, patch_size=16, embed_dim=64, depth=2, num_heads=1, vocab_size=args.vocab_size,
            max_text_len=args.seq_len, drop_path_rate=args.dropout, shared_param=args.shared_param,
            share_scope=args.share_scope, colearn_param=args.colearn_param, **kwargs)
        model.sync_shared_weights()
        return model

    @register_model
    def mome_base_patch16(pretrained, args, **kwargs):
        """ViT-B sized (the shipped mome_base_patch16_224_ours factory is broken, mome.py:1009)."""
        model = mome.ModalityAgnosticTransformer(
            img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, vocab_size=args.vocab_size,
            max_text_len=args.seq_len, drop_path_rate=args.dropout, shared_param=args.shared_param,
            share_scope=args.share_scope, colearn_param=args.colearn_param, **kwargs)
        model.sync_shared_weights()
        return model

    _installed = True


class NullWriter:
    def log(self, *a, **k):
        pass

    def finish(self):
        pass
