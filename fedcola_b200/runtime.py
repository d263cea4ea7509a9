"""Python side of the native step driver (csrc/mat_driver.cu): descriptor/table construction, device
buffers, the autograd bridge used by `ModalityAgnosticTransformer.forward`, and the fused `ClientTrainer`
used by the client update loop.  Everything here is plumbing — tensors are allocated by torch, every
kernel is launched by ONE ctypes call into libfedcola_b200.so on torch's current stream."""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .arena import BLOCK_ROLES, MatSpec

MAX_DEPTH, MAX_SEGMENTS, N_ROLES = 24, 2048, 20
LOSS_CE_IMG, LOSS_CE_TXT, LOSS_CONTRASTIVE = 0, 1, 2
OPT_NONE, OPT_ADAMW, OPT_SGD = 0, 1, 2
LIN_SHORT = ("qkv", "proj", "fc1", "fc2")

c_ll, c_int, c_float, c_vp = ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_void_p


class MatDesc(ctypes.Structure):
    _fields_ = [("d", c_int), ("depth", c_int), ("heads", c_int), ("hidden", c_int),
                ("img_size", c_int), ("patches", c_int), ("in_chans", c_int),
                ("seq_len", c_int), ("vocab", c_int), ("max_text_len", c_int),
                ("num_classes", c_int * 2), ("has_enc", c_int * 2), ("with_aux", c_int), ("aux_trained", c_int),
                ("precise", c_int), ("op_lo_offset", c_ll),
                ("img_pos", c_ll), ("img_cls", c_ll), ("img_pw", c_ll), ("img_pb", c_ll),
                ("txt_word", c_ll), ("txt_pos", c_ll), ("txt_type", c_ll), ("txt_lnw", c_ll), ("txt_lnb", c_ll),
                ("norm_w", c_ll), ("norm_b", c_ll), ("head_w", c_ll * 2), ("head_b", c_ll * 2),
                ("blk", ((c_ll * N_ROLES) * MAX_DEPTH) * 2),
                ("op_pw", c_ll), ("op", ((c_ll * 4) * MAX_DEPTH) * 2)]


class StepArgs(ctypes.Structure):
    _fields_ = [("B", c_int), ("loss_kind", c_int), ("optimizer", c_int), ("step", c_int),
                ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("weight_decay", c_float),
                ("momentum", c_float), ("dampening", c_float), ("nesterov", c_int),
                ("max_grad_norm", c_float), ("prox_mu", c_float),
                ("params", c_vp), ("grads", c_vp), ("opt_state0", c_vp), ("opt_state1", c_vp),
                ("global_params", c_vp), ("operands", c_vp), ("workspace", c_vp), ("arena_floats", c_ll),
                ("img", c_vp), ("ids", c_vp), ("labels", c_vp), ("droppath", c_vp), ("stats", c_vp),
                ("chunks", c_vp), ("n_chunks", c_int), ("n_segments", c_int),
                ("prep_layers", c_vp), ("n_prep_layers", c_int), ("n_prep_tiles", c_int),
                ("aux_layers", c_vp), ("n_aux_layers", c_int), ("n_aux_chunks", c_int),
                ("fused_a", c_vp), ("n_fused_a", c_int), ("fused_b", c_vp), ("n_fused_b", c_int),
                ("fused_counters", c_vp)]


CHUNK_DT = np.dtype([("off", "<i8"), ("len", "<i4"), ("seg", "<i4")], align=True)
PREP_DT = np.dtype([("w_off", "<i8"), ("a_off", "<i8"), ("s_off", "<i8"), ("dst_off", "<i8"), ("dstT_off", "<i8"),
                    ("rows", "<i4"), ("cols", "<i4"), ("tile_start", "<i4")], align=True)
FCHUNK_DT = np.dtype([("off", "<i8"), ("len", "<i4"), ("kind", "<i4"), ("layer", "<i4"), ("pad", "<i4"), ("x0", "<i8"),
                      ("x1", "<i8")], align=True)
AUX_DT = np.dtype([("w_off", "<i8"), ("a_off", "<i8"), ("s_off", "<i8"), ("numel", "<i8"), ("chunk_start", "<i4")],
                  align=True)


def _round_up(x, a):
    return (x + a - 1) // a * a


class ModelPlan:
    """Everything static about one MatSpec: C descriptor, operand-arena layout, prep/aux tables."""

    def __init__(self, spec: MatSpec, precise=False):
        _sigs()
        self.precise = bool(precise)
        if spec.depth > MAX_DEPTH:
            raise ValueError(f"depth {spec.depth} > {MAX_DEPTH}")
        if spec.head_dim != 64:
            raise NotImplementedError("the sm_100a attention kernel supports head_dim == 64 (all reference factories)")
        self.spec = spec
        m = MatDesc()
        m.d, m.depth, m.heads, m.hidden = spec.embed_dim, spec.depth, spec.num_heads, spec.embed_dim * spec.mlp_ratio
        m.img_size, m.patches, m.in_chans = spec.img_size, spec.num_patches, spec.in_chans
        m.seq_len, m.vocab, m.max_text_len = spec.max_text_len, spec.vocab_size, spec.max_text_len
        m.with_aux, m.aux_trained = int(spec.has_aux), int(spec.aux_trained)
        r = spec.role
        for e in range(2):
            m.has_enc[e] = int(spec.modalities[e] is not None)
            nc = spec.num_classes[e]
            m.num_classes[e] = int(nc) if (nc and spec.tasks[e] == "cls") else 0
            m.head_w[e], m.head_b[e] = r(f"head.{e}.w"), r(f"head.{e}.b")
        m.img_pos, m.img_cls, m.img_pw, m.img_pb = r("img.pos"), r("img.cls"), r("img.pw"), r("img.pb")
        m.txt_word, m.txt_pos, m.txt_type = r("txt.word"), r("txt.pos"), r("txt.type")
        m.txt_lnw, m.txt_lnb = r("txt.lnw"), r("txt.lnb")
        m.norm_w, m.norm_b = r("norm.w"), r("norm.b")
        prep, aux, off = [], [], 0
        d, hid = m.d, m.hidden
        m.op_pw = -1
        if m.has_enc[0]:
            m.op_pw = off
            prep.append((m.img_pw, -1, -1, off, d, 768))
            off = _round_up(off + d * 768, 64)
        shapes = {"qkv": (3 * d, d), "proj": (d, d), "fc1": (hid, d), "fc2": (d, hid)}
        for e in range(2):
            for j in range(MAX_DEPTH):
                for k in range(N_ROLES):
                    m.blk[e][j][k] = r(f"blk.{e}.{j}.{BLOCK_ROLES[k]}") if (m.has_enc[e] and j < spec.depth) else -1
                for li, ln in enumerate(LIN_SHORT):
                    m.op[e][j][li] = -1
                    if m.has_enc[e] and j < spec.depth:
                        rows, cols = shapes[ln]
                        w, a, s = r(f"blk.{e}.{j}.{ln}w"), r(f"blk.{e}.{j}.{ln}a"), r(f"blk.{e}.{j}.{ln}s")
                        m.op[e][j][li] = off
                        prep.append((w, a, s, off, rows, cols))
                        if a >= 0:
                            aux.append((w, a, s, rows * cols))
                        off = _round_up(off + rows * cols, 64)
        self.desc = m
        self.operand_elems = max(off, 64)
        # fp32-accurate validation mode: the operand arena holds W_eff as (hi, lo) bf16 pairs — lo parts after the hi parts
        m.precise = int(self.precise)
        m.op_lo_offset = _round_up(self.operand_elems, 64) if self.precise else 0
        self.operand_alloc = m.op_lo_offset + self.operand_elems if self.precise else self.operand_elems
        pt = np.zeros(len(prep), dtype=PREP_DT)
        tiles = 0
        for i, (w, a, s, dst, rows, cols) in enumerate(prep):
            pt[i] = (w, a, s, dst, -1, rows, cols, tiles)
            tiles += ((rows + 31) // 32) * ((cols + 31) // 32)
        self.prep_table, self.n_prep_tiles = pt, tiles
        chunk = int(_lib.lib().fc_chunk_floats())
        at = np.zeros(len(aux), dtype=AUX_DT)
        nch = 0
        for i, (w, a, s, numel) in enumerate(aux):
            at[i] = (w, a, s, numel, nch)
            nch += (numel + chunk - 1) // chunk
        self.aux_table, self.n_aux_chunks = at, nch
        self.chunk = chunk

    def chunk_table(self, requires_grad):
        """Work items over the trainable segments. requires_grad: {key: bool}. Returns (table, n_segments).
        Vectorised and cached per requires_grad signature (a ViT-S client has ~21k chunks)."""
        segs = [s for s in self.spec.unique_segments() if requires_grad.get(s.key, s.requires_grad)]
        sig = tuple(s.key for s in segs)
        cache = self.__dict__.setdefault("_chunk_cache", {})
        if sig in cache:
            return cache[sig]
        if len(segs) > MAX_SEGMENTS:
            raise ValueError("too many parameter tensors")
        offs = np.asarray([s.offset for s in segs], dtype=np.int64)
        nums = np.asarray([s.numel for s in segs], dtype=np.int64)
        per = (nums + self.chunk - 1) // self.chunk
        seg_id = np.repeat(np.arange(len(segs), dtype=np.int64), per)
        first = np.concatenate([[0], np.cumsum(per)[:-1]])
        k = np.arange(int(per.sum()), dtype=np.int64) - np.repeat(first, per)      # chunk index inside its segment
        t = np.zeros(len(k), dtype=CHUNK_DT)
        t["off"] = offs[seg_id] + k * self.chunk
        t["len"] = np.minimum(self.chunk, nums[seg_id] - k * self.chunk)
        t["seg"] = seg_id
        cache[sig] = (t, len(segs))
        return cache[sig]

    def fused_tables(self, requires_grad):
        """Chunk tables of the fused optimizer tail (fc_opt_fused): (phase A, phase B) or None when the combination of
        frozen / trainable tensors is not one it covers (then the separate aux_grads / step / prep kernels run).
        Cached per requires_grad signature."""
        segs = [s for s in self.spec.unique_segments() if requires_grad.get(s.key, s.requires_grad)]
        sig = tuple(s.key for s in segs)
        cache = self.__dict__.setdefault("_fused_cache", {})
        if sig in cache:
            return cache[sig]
        by_off = {s.offset: s for s in segs}
        lin = {}          # W offset -> (operand offset, A offset, s offset, aux layer index)
        aux_index = {int(r["w_off"]): i for i, r in enumerate(self.aux_table)}
        for r in self.prep_table:
            lin[int(r["w_off"])] = (int(r["dst_off"]), int(r["a_off"]), int(r["s_off"]), aux_index.get(int(r["w_off"]), -1))
        a_of, s_of, ok = {}, set(), True
        for w, (dst, a, s_, li) in lin.items():
            if a >= 0:
                # an aux layer is covered when W, aux_weight and cross_modal_scale all train (--aux_trained)
                if not (w in by_off and a in by_off and s_ in by_off and li >= 0):
                    ok = False
                a_of[a] = (w, li)
                s_of.add(s_)
        if not ok:
            cache[sig] = None
            return None
        rows_a, rows_b = [], []

        def chunks(seg):
            for k in range(0, seg.numel, self.chunk):
                yield k, min(self.chunk, seg.numel - k)
        for sg in segs:
            if sg.offset in s_of:
                continue                                   # stepped by the block that completes its layer's <dW, A>
            if sg.offset in lin:
                dst, a, s_, li = lin[sg.offset]
                for k, n in chunks(sg):
                    rows_b.append((sg.offset + k, n, 3 if a >= 0 else 2, max(li, 0), 0, dst + k, a + k if a >= 0 else 0))
            elif sg.offset in a_of:
                w, li = a_of[sg.offset]
                for k, n in chunks(sg):
                    rows_a.append((sg.offset + k, n, 1, li, 0, w + k, 0))
            else:
                for k, n in chunks(sg):
                    rows_a.append((sg.offset + k, n, 0, 0, 0, 0, 0))
        ta = np.array(rows_a, dtype=FCHUNK_DT) if rows_a else np.zeros(0, dtype=FCHUNK_DT)
        tb = np.array(rows_b, dtype=FCHUNK_DT) if rows_b else np.zeros(0, dtype=FCHUNK_DT)
        cache[sig] = (ta, tb)
        return cache[sig]

    def workspace_bytes(self, B):
        n = _lib.lib().fc_mat_workspace_bytes(ctypes.byref(self.desc), int(B))
        if n < 0:
            _lib.check(-1, "fc_mat_workspace_bytes")
        return int(n)


_lib_sigs_done = False


def _sigs():
    global _lib_sigs_done
    L = _lib.lib()
    if not _lib_sigs_done:
        L.fc_mat_workspace_bytes.restype = ctypes.c_longlong
        if L.fc_sizeof_mat_desc() != ctypes.sizeof(MatDesc) or L.fc_sizeof_step_args() != ctypes.sizeof(StepArgs):
            raise RuntimeError("fedcola_b200: ctypes struct layout does not match include/fedcola_b200.h")
        _lib_sigs_done = True
    return L


def _to_dev(np_table, device):
    if np_table.size == 0:
        return None
    t = torch.from_numpy(np_table.view(np.uint8).reshape(-1).copy())
    return t.to(device)


_PLAN_CACHE = {}
_PLAN_LOCK = __import__("threading").Lock()      # client worker threads build their first runtime concurrently


def plan_for(spec: MatSpec, precise=False) -> ModelPlan:
    """ModelPlans are static per architecture: cache them (a round creates one runtime per sampled client)."""
    sig = (bool(precise), spec.embed_dim, spec.depth, spec.num_heads, spec.modalities, spec.num_classes, spec.tasks, spec.vocab_size,
           spec.max_text_len, spec.img_size, spec.patch_size, spec.in_chans, spec.mlp_ratio, spec.with_aux,
           spec.aux_trained, spec.aux_attn_only, spec.aux_mlp_only, spec.share_scope, tuple(spec.keys()))
    p = _PLAN_CACHE.get(sig)
    if p is None:
        with _PLAN_LOCK:
            p = _PLAN_CACHE.get(sig)
            if p is None:
                p = _PLAN_CACHE[sig] = ModelPlan(spec, precise)
    return p


class ModelRuntime:
    """Device buffers of one model instance: bf16 operand arena, workspace, grad arena, tables."""

    def __init__(self, spec: MatSpec, arena: torch.Tensor, precise=False):
        _lib.require_cuda(arena, "model arena")
        _sigs()
        self.plan = plan_for(spec, precise)
        self.arena = arena
        self.device = arena.device
        self.dev_index = arena.device.index if arena.device.index is not None else torch.cuda.current_device()
        self.operands = torch.zeros(self.plan.operand_alloc, dtype=torch.bfloat16, device=self.device)
        dev_tables = self.plan.__dict__.setdefault("_dev_tables", {})
        if str(self.device) not in dev_tables:
            dev_tables[str(self.device)] = (_to_dev(self.plan.prep_table, self.device),
                                            _to_dev(self.plan.aux_table, self.device))
        self.prep_dev, self.aux_dev = dev_tables[str(self.device)]
        self.workspace, self.ws_batch = None, 0
        self.grads = None
        self.operands_version = None

    def ensure_workspace(self, B):
        if self.workspace is None or self.ws_batch < B:
            self.workspace = torch.empty(self.plan.workspace_bytes(B), dtype=torch.uint8, device=self.device)
            self.ws_batch = B
        elif self.ws_batch != B:
            # layouts depend on B; a smaller batch (last partial batch) re-carves inside the same allocation
            pass
        return self.workspace

    def ensure_grads(self):
        if self.grads is None:
            self.grads = torch.zeros_like(self.arena)
        return self.grads

    def refresh_operands(self):
        """bf16 W_eff = W + s*A for every Linear (fc_prep_weights; as (hi, lo) pairs in the validation mode)."""
        p = self.plan
        if p.precise:
            lo = self.operands[p.desc.op_lo_offset:]
            rc = _lib.lib().fc_prep_weights_split(_lib.ptr(self.arena), _lib.ptr(self.operands), _lib.ptr(lo),
                                                  _lib.ptr(self.prep_dev), c_int(len(p.prep_table)), c_int(self.dev_index),
                                                  _lib.stream_ptr(self.device))
            _lib.check(rc, "fc_prep_weights_split")
            return
        rc = _lib.lib().fc_prep_weights(_lib.ptr(self.arena), _lib.ptr(self.operands), _lib.ptr(self.prep_dev),
                                        c_int(len(p.prep_table)), c_int(p.n_prep_tiles), c_int(self.dev_index),
                                        _lib.stream_ptr(self.device))
        _lib.check(rc, "fc_prep_weights")

    def forward(self, B, img, ids, droppath, feat_out, out0, out1):
        rc = _lib.lib().fc_mat_forward(ctypes.byref(self.plan.desc), _lib.ptr(self.arena), _lib.ptr(self.operands),
                                       _lib.ptr(self.ensure_workspace(B)), c_int(B), _lib.ptr(img), _lib.ptr(ids),
                                       _lib.ptr(droppath), c_int(int(feat_out)), _lib.ptr(out0), _lib.ptr(out1),
                                       c_int(self.dev_index), _lib.stream_ptr(self.device))
        _lib.check(rc, "fc_mat_forward")

    def backward(self, B, ids, droppath, feat_out, dout0, dout1, grads):
        p = self.plan
        rc = _lib.lib().fc_mat_backward(ctypes.byref(p.desc), _lib.ptr(self.arena), _lib.ptr(self.operands),
                                        _lib.ptr(self.workspace), c_int(B), _lib.ptr(ids), _lib.ptr(droppath),
                                        c_int(int(feat_out)), _lib.ptr(dout0), _lib.ptr(dout1), _lib.ptr(grads),
                                        _lib.ptr(self.aux_dev), c_int(len(p.aux_table)), c_int(p.n_aux_chunks),
                                        c_int(self.dev_index), _lib.stream_ptr(self.device))
        _lib.check(rc, "fc_mat_backward")


def droppath_scales(spec: MatSpec, B, device, training, mode="reference"):
    """fp32 [2, depth, 2, B] per-sample scales (mask/keep) or None.  mode='reference' draws from torch's
    generator with the reference's call sequence (timm DropPath: new_empty((B,1,1)).bernoulli_(keep), one
    call per block branch, encoder 0 first — mome.py:226-227); mode='fused' draws all masks in one call."""
    rate = spec.drop_path_rate
    if not training or rate <= 0.0:
        return None
    dpr = [x.item() for x in torch.linspace(0, rate, spec.depth)]     # mome.py:726-728
    out = torch.ones(2, spec.depth, 2, B, dtype=torch.float32, device=device)
    if mode == "fused":
        keep = torch.tensor([1.0 - r for r in dpr], dtype=torch.float32, device=device).view(1, -1, 1, 1)
        m = torch.bernoulli(keep.expand(2, spec.depth, 2, B))
        return (m / keep).contiguous()
    for e in range(2):
        if spec.modalities[e] is None:
            continue
        for j, r in enumerate(dpr):
            if r <= 0.0:
                continue
            keep = 1.0 - r
            for br in range(2):
                m = torch.empty((B, 1, 1), dtype=torch.float32, device=device).bernoulli_(keep)
                if keep > 0.0:
                    m.div_(keep)
                out[e, j, br] = m.view(B)
    return out


def _runtime(model):
    rt = model._runtime
    precise = getattr(model, "precision", "bf16") == "fp32"
    if rt is None or rt.arena.data_ptr() != model.arena.data_ptr() or rt.plan.precise != precise:
        rt = ModelRuntime(model.spec, model.arena, precise)
        model._runtime = rt
    return rt


class _MatFunction(torch.autograd.Function):
    """Autograd bridge: forward/backward of the whole model in two native calls."""

    @staticmethod
    def forward(ctx, model, img, ids, feat_out, droppath, *params):
        rt = _runtime(model)
        spec = model.spec
        B = (img if img is not None else ids).shape[0]
        rt.refresh_operands()
        outs = []
        for e in range(2):
            if spec.modalities[e] is None:
                outs.append(None)
            else:
                C = rt.plan.desc.num_classes[e]
                width = spec.embed_dim if (feat_out or C <= 0) else C
                outs.append(torch.empty(B, width, dtype=torch.float32, device=rt.device))
        rt.forward(B, img, ids, droppath, feat_out, outs[0], outs[1])
        ctx.model, ctx.rt, ctx.B, ctx.ids, ctx.feat_out, ctx.droppath = model, rt, B, ids, feat_out, droppath
        ctx.present = [o is not None for o in outs]
        ret = tuple(o if o is not None else torch.empty(0, device=rt.device) for o in outs)
        ctx.mark_non_differentiable(*[r for r, p in zip(ret, ctx.present) if not p])
        return ret

    @staticmethod
    def backward(ctx, g0, g1):
        rt, model = ctx.rt, ctx.model
        grads = rt.ensure_grads()
        grads.zero_()
        d = [g.contiguous().float() if (g is not None and p) else None for g, p in zip((g0, g1), ctx.present)]
        rt.backward(ctx.B, ctx.ids, ctx.droppath, ctx.feat_out, d[0], d[1], grads)
        out = []
        for seg in model.spec.unique_segments():
            p = model._params_by_key[seg.key]
            out.append(grads[seg.offset:seg.offset + seg.numel].view(seg.shape) if p.requires_grad else None)
        return (None, None, None, None, None, *out)


def model_forward(model, x, feat_out=False):
    """ModalityAgnosticTransformer.forward (mome.py:881-922) on the native driver."""
    spec = model.spec
    if not model.arena.is_cuda:
        raise RuntimeError("fedcola_b200: the model must live on a CUDA device (there is no CPU fallback); "
                           "call model.to('cuda') first")
    img = ids = None
    for i, m in enumerate(spec.modalities):
        if m is None:
            assert x[i] is None, "None modality should have None input."
        elif m == "img":
            img = x[i].to(model.device, torch.float32).contiguous()
        else:
            ids = x[i].to(model.device, torch.int64).contiguous()
    B = (img if img is not None else ids).shape[0]
    dp = droppath_scales(spec, B, model.device, model.training, getattr(model, "droppath_rng", "reference"))
    params = [model._params_by_key[s.key] for s in spec.unique_segments()]
    o0, o1 = _MatFunction.apply(model, img, ids, bool(feat_out), dp, *params)
    outs = [None, None]
    for i, (m, o) in enumerate(zip(spec.modalities, (o0, o1))):
        if m is not None:
            outs[i] = o
    return outs


class ClientTrainer:
    """Fused local training of one client model: one native call per batch (fc_client_step)."""

    def __init__(self, model, optimizer="AdamW", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 momentum=0.0, dampening=0.0, nesterov=False, max_grad_norm=0.0, prox_mu=0.0, global_arena=None):
        self.model = model
        self.rt = _runtime(model)
        rt = self.rt
        if optimizer == "AdamW":
            self.opt = OPT_ADAMW
        elif optimizer == "SGD":
            self.opt = OPT_SGD
        else:
            raise NotImplementedError(f"fedcola_b200: optimizer {optimizer!r} has no fused sm_100a step "
                                      "(supported: AdamW, SGD)")
        flags = {k: p.requires_grad for k, p in model._params_by_key.items()}
        table, nseg = rt.plan.chunk_table(flags)
        self.n_chunks, self.n_segments = len(table), nseg
        cdev = rt.plan.__dict__.setdefault("_chunk_dev", {})
        ck = (id(table), str(rt.device))
        if ck not in cdev:
            cdev[ck] = _to_dev(table, rt.device)
        self.chunks_dev = cdev[ck]
        self.grads = rt.ensure_grads()
        # fresh optimizer state every round, as the reference re-creates the optimizer (fedavgclient.py:63)
        self.state0 = torch.zeros_like(model.arena) if (self.opt == OPT_ADAMW or momentum != 0.0) else None
        self.state1 = torch.zeros_like(model.arena) if self.opt == OPT_ADAMW else None
        self.stats = torch.zeros(4, dtype=torch.float32, device=rt.device)
        # fused optimizer tail (aux gradients + step + bf16 operand refresh in two launches), when it applies
        self.fused = None
        ft = rt.plan.fused_tables(flags) if (not rt.plan.precise and max_grad_norm <= 0 and prox_mu <= 0) else None
        if ft is not None:
            fk = ("fused", id(ft), str(rt.device))
            if fk not in cdev:
                cdev[fk] = (_to_dev(ft[0], rt.device), _to_dev(ft[1], rt.device), len(ft[0]), len(ft[1]))
            self.fused = cdev[fk]
            self.fused_counters = torch.zeros(max(len(rt.plan.aux_table), 1), dtype=torch.int32, device=rt.device)
        self.step_count = 0
        self.global_arena = global_arena
        a = StepArgs()
        a.optimizer, a.lr, a.beta1, a.beta2, a.eps = self.opt, lr, betas[0], betas[1], eps
        a.weight_decay, a.momentum, a.dampening, a.nesterov = weight_decay, momentum, dampening, int(nesterov)
        a.max_grad_norm, a.prox_mu = max_grad_norm, prox_mu
        self.args = a
        rt.refresh_operands()

    def prepare(self, img, ids, labels, loss_kind, droppath=None):
        """Fill the native argument block for one batch (no launch). Returns the batch size."""
        rt, a, p = self.rt, self.args, self.rt.plan
        for t, dt, name in ((img, torch.float32, "images"), (ids, torch.int64, "token ids"), (labels, torch.int64, "labels")):
            if t is not None and (t.dtype != dt or not t.is_contiguous() or not t.is_cuda):
                raise TypeError(f"fedcola_b200: {name} must be a contiguous CUDA {dt} tensor (got {t.dtype}, "
                                f"contiguous={t.is_contiguous()}, device={t.device}): the kernels read raw pointers")
        B = (img if img is not None else ids).shape[0]
        self.step_count += 1
        a.B, a.loss_kind, a.step = B, loss_kind, self.step_count
        a.params, a.grads = self.model.arena.data_ptr(), self.grads.data_ptr()
        a.opt_state0 = self.state0.data_ptr() if self.state0 is not None else None
        a.opt_state1 = self.state1.data_ptr() if self.state1 is not None else None
        a.global_params = self.global_arena.data_ptr() if self.global_arena is not None else None
        a.operands, a.workspace = rt.operands.data_ptr(), rt.ensure_workspace(B).data_ptr()
        a.arena_floats = self.model.arena.numel()
        a.img = img.data_ptr() if img is not None else None
        a.ids = ids.data_ptr() if ids is not None else None
        a.labels = labels.data_ptr() if labels is not None else None
        a.droppath = droppath.data_ptr() if droppath is not None else None
        a.stats = self.stats.data_ptr()
        a.chunks, a.n_chunks, a.n_segments = (self.chunks_dev.data_ptr() if self.chunks_dev is not None else None,
                                              self.n_chunks, self.n_segments)
        a.prep_layers = rt.prep_dev.data_ptr() if rt.prep_dev is not None else None
        a.n_prep_layers, a.n_prep_tiles = len(p.prep_table), p.n_prep_tiles
        a.aux_layers = rt.aux_dev.data_ptr() if rt.aux_dev is not None else None
        a.n_aux_layers, a.n_aux_chunks = len(p.aux_table), p.n_aux_chunks
        if self.fused is not None:
            fa, fb, na, nb = self.fused
            a.fused_a, a.n_fused_a = (fa.data_ptr() if fa is not None else None), na
            a.fused_b, a.n_fused_b = (fb.data_ptr() if fb is not None else None), nb
            a.fused_counters = self.fused_counters.data_ptr()
        else:
            a.fused_a = a.fused_b = a.fused_counters = None
            a.n_fused_a = a.n_fused_b = 0
        self._keep = (img, ids, labels, droppath)      # the launch is asynchronous: keep the batch alive
        return B

    def step(self, img, ids, labels, loss_kind, droppath=None):
        self.prepare(img, ids, labels, loss_kind, droppath)
        group_step([self])


MAX_GROUP = 4          # FC_STEP_MAX_GROUPS


def group_step(trainers):
    """ONE native call trains one batch of every trainer of a lockstep group (fc_client_step_group): all trainers hold
    the same model architecture on the same device and were `prepare`d with batches of the same size."""
    t0 = trainers[0]
    if not 1 <= len(trainers) <= MAX_GROUP:
        raise ValueError(f"a lockstep group holds 1..{MAX_GROUP} clients")
    for t in trainers[1:]:
        if t.rt.plan is not t0.rt.plan or t.rt.device != t0.rt.device or t.args.B != t0.args.B:
            raise ValueError("clients of a lockstep group must share the model architecture, the device and the batch size")
    arr = (ctypes.POINTER(StepArgs) * len(trainers))(*[ctypes.pointer(t.args) for t in trainers])
    rc = _lib.lib().fc_client_step_group(ctypes.byref(t0.rt.plan.desc), c_int(len(trainers)), arr, c_int(t0.rt.dev_index),
                                         _lib.stream_ptr(t0.rt.device))
    _lib.check(rc, "fc_client_step_group")
