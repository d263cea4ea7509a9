"""Flat fp32 parameter arenas keyed by the reference's state_dict names.

Every model kind of the reference (`ModalityAgnosticTransformer` with modalities [img,None], [None,txt]
or [img,txt]; /root/reference/src/models/mome.py:671-786) is laid out as ONE contiguous fp32 buffer in
HBM.  `MatSpec.segments` lists, in the reference's `state_dict()` order, where each tensor lives; keys,
shapes and ordering are checked against the real reference in tests/test_oracle_vs_reference.py.

Layout rules
  * every tensor starts on a 128-byte boundary (32 floats) so 128-bit vector loads and TMA stay aligned;
  * `share_scope == 'all'` aliases (mome.py:824-827: `blockses[None-idx]` *is* the main encoder) are extra
    keys that point at the same offsets (`Segment.alias_of`);
  * `aux_weight` / `cross_modal_scale` (CrossModalReparamLinear, mome.py:42-60) live in the arena right
    after their layer's bias, in the reference's registration order (weight, bias, scale, aux).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

ALIGN = 32  # floats (128 B)

AUX_LAYERS_ALL = ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")


def _round_up(x, a=ALIGN):
    return (x + a - 1) // a * a


@dataclass
class Segment:
    key: str
    shape: Tuple[int, ...]
    offset: int            # in floats
    numel: int
    requires_grad: bool = True
    alias_of: Optional[str] = None
    role: str = ""         # e.g. 'blk.qkv.weight' — used by the native step driver's offset table


# roles of the per-block tensors, in arena order; the native driver indexes its offset table by these
BLOCK_ROLES = ("n1w", "n1b", "qkvw", "qkvb", "qkvs", "qkva", "projw", "projb", "projs", "proja",
               "n2w", "n2b", "fc1w", "fc1b", "fc1s", "fc1a", "fc2w", "fc2b", "fc2s", "fc2a")


@dataclass
class MatSpec:
    """Static description of one ModalityAgnosticTransformer instance."""
    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    modalities: Tuple[Optional[str], Optional[str]] = ("img", None)
    num_classes: Tuple[Optional[int], Optional[int]] = (100, None)
    tasks: Tuple[Optional[str], Optional[str]] = ("cls", None)
    vocab_size: int = 30522
    max_text_len: int = 40
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    mlp_ratio: int = 4
    drop_path_rate: float = 0.0
    with_aux: bool = False
    aux_trained: bool = False
    aux_attn_only: bool = False
    aux_mlp_only: bool = False
    share_scope: str = "dataset"
    shared_param: str = "none"
    colearn_param: str = "none"
    segments: List[Segment] = field(default_factory=list, repr=False)
    total: int = 0

    def __post_init__(self):
        self.modalities = tuple(self.modalities)
        self.num_classes = tuple(self.num_classes)
        self.tasks = tuple(self.tasks)
        if self.embed_dim % self.num_heads:
            raise ValueError("embed_dim must be divisible by num_heads")
        if self.colearn_param == "attn" and None not in self.modalities:
            # mome.py:837-841 shares attn modules between encoders; not used by any BASELINE config.
            raise NotImplementedError("colearn_param='attn' is not supported by the B200 path yet")
        self._build()

    # ---- derived -------------------------------------------------------------------------------
    @property
    def signature(self):
        """Hashable identity of the layout: equal for deep copies (the segment table is a pure function of it)."""
        return (self.embed_dim, self.depth, self.num_heads, self.modalities, self.num_classes, self.tasks,
                self.vocab_size, self.max_text_len, self.img_size, self.patch_size, self.in_chans, self.mlp_ratio,
                self.with_aux, self.aux_trained, self.aux_attn_only, self.aux_mlp_only, self.share_scope,
                self.shared_param, self.colearn_param, self.total)

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads

    @property
    def num_patches(self):
        return (self.img_size // self.patch_size) ** 2

    @property
    def has_aux(self):
        # build_aux only runs for uni-modal models (mome.py:768-769)
        return self.with_aux and (None in self.modalities)

    @property
    def main_idx(self):
        return 0 if self.modalities[0] is not None else 1

    def aux_layer_names(self):
        if self.aux_attn_only:
            if self.aux_mlp_only:
                raise ValueError("Both aux_attn_only and aux_mlp_only cannot be True.")
            return ("attn.qkv", "attn.proj")
        if self.aux_mlp_only:
            return ("mlp.fc1", "mlp.fc2")
        return AUX_LAYERS_ALL

    # ---- layout --------------------------------------------------------------------------------
    def _add(self, key, shape, role="", requires_grad=True):
        n = 1
        for s in shape:
            n *= s
        seg = Segment(key, tuple(shape), self.total, n, requires_grad, None, role)
        self.segments.append(seg)
        self.total = _round_up(self.total + n)
        return seg

    def _build(self):
        d, hid = self.embed_dim, self.embed_dim * self.mlp_ratio
        self.segments, self.total = [], 0
        aux_names = self.aux_layer_names() if self.has_aux else ()
        for i, m in enumerate(self.modalities):
            if m == "img":
                p = f"embeddings.{i}."
                self._add(p + "pos_embed", (1, self.num_patches + 1, d), "img.pos")
                self._add(p + "cls_token", (1, 1, d), "img.cls")
                self._add(p + "embed.proj.weight", (d, self.in_chans, self.patch_size, self.patch_size), "img.pw")
                self._add(p + "embed.proj.bias", (d,), "img.pb")
            elif m == "txt":
                p = f"embeddings.{i}.text_embeddings."
                self._add(p + "word_embeddings.weight", (self.vocab_size, d), "txt.word")
                self._add(p + "position_embeddings.weight", (self.max_text_len, d), "txt.pos")
                self._add(p + "token_type_embeddings.weight", (2, d), "txt.type")
                self._add(p + "LayerNorm.weight", (d,), "txt.lnw")
                self._add(p + "LayerNorm.bias", (d,), "txt.lnb")
        block_keys = {}
        for i, m in enumerate(self.modalities):
            if m is None:
                continue
            keys = []
            for j in range(self.depth):
                p = f"blockses.{i}.{j}."
                r = f"blk.{i}.{j}."
                keys.append(self._add(p + "norm1.weight", (d,), r + "n1w"))
                keys.append(self._add(p + "norm1.bias", (d,), r + "n1b"))
                for lname, short, shp in (("attn.qkv", "qkv", (3 * d, d)), ("attn.proj", "proj", (d, d))):
                    keys.append(self._add(p + lname + ".weight", shp, r + short + "w"))
                    keys.append(self._add(p + lname + ".bias", (shp[0],), r + short + "b"))
                    if lname in aux_names:
                        keys.append(self._add(p + lname + ".cross_modal_scale", (1,), r + short + "s"))
                        keys.append(self._add(p + lname + ".aux_weight", shp, r + short + "a", self.aux_trained))
                keys.append(self._add(p + "norm2.weight", (d,), r + "n2w"))
                keys.append(self._add(p + "norm2.bias", (d,), r + "n2b"))
                for lname, short, shp in (("mlp.fc1", "fc1", (hid, d)), ("mlp.fc2", "fc2", (d, hid))):
                    keys.append(self._add(p + lname + ".weight", shp, r + short + "w"))
                    keys.append(self._add(p + lname + ".bias", (shp[0],), r + short + "b"))
                    if lname in aux_names:
                        keys.append(self._add(p + lname + ".cross_modal_scale", (1,), r + short + "s"))
                        keys.append(self._add(p + lname + ".aux_weight", shp, r + short + "a", self.aux_trained))
            block_keys[i] = keys
        # share_scope == 'all': the None encoder aliases the main one (state_dict lists both prefixes,
        # blockses.0.* first).  Aliases are inserted in state_dict order below.
        if self.share_scope == "all" and None in self.modalities:
            main, none_idx = self.main_idx, 1 - self.main_idx
            alias = [Segment(s.key.replace(f"blockses.{main}.", f"blockses.{none_idx}.", 1), s.shape, s.offset,
                             s.numel, s.requires_grad, s.key, "") for s in block_keys[main]]
            first_blk = self.segments.index(block_keys[main][0])
            if none_idx < main:   # alias keys come first in state_dict order
                self.segments[first_blk:first_blk] = alias
            else:
                self.segments.extend(alias)
        self._add("norm.weight", (d,), "norm.w")
        self._add("norm.bias", (d,), "norm.b")
        for i, t in enumerate(self.tasks):
            if t == "cls" and self.num_classes[i] and self.num_classes[i] > 0:
                self._add(f"heads.{i}.head.weight", (self.num_classes[i], d), f"head.{i}.w")
                self._add(f"heads.{i}.head.bias", (self.num_classes[i],), f"head.{i}.b")
        self._by_key = {s.key: s for s in self.segments}
        self._by_role = {s.role: s for s in self.segments if s.role}

    # ---- queries -------------------------------------------------------------------------------
    def keys(self):
        return [s.key for s in self.segments]

    def seg(self, key):
        return self._by_key[key]

    def role(self, role):
        s = self._by_role.get(role)
        return s.offset if s is not None else -1

    def unique_segments(self):
        return [s for s in self.segments if s.alias_of is None]

    def required_keys(self):
        """Keys of `required_params()` (mome.py:844-860): drop blocks of None encoders and aux/scale keys."""
        out = []
        for s in self.segments:
            if any(m is None and f"blockses.{i}" in s.key for i, m in enumerate(self.modalities)):
                continue
            if self.with_aux and ("aux" in s.key or "cross_modal_scale" in s.key):
                continue
            out.append(s.key)
        return out

    def aux_keys(self):
        """Keys of `aux_params()` (mome.py:862-878)."""
        none_idx = [i for i, m in enumerate(self.modalities) if m is None]
        return [s.key for s in self.segments
                if "aux" in s.key and not any(f"blockses.{i}" in s.key for i in none_idx)]

    def n_params(self):
        return sum(s.numel for s in self.unique_segments())
