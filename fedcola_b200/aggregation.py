"""Host-side planner for the streaming aggregation kernel (csrc/aggregate.cu).

Mirrors the coefficient stage of FedavgServer._aggregate (/root/reference/src/server/fedavgserver.py:
601-653) in Python — bookkeeping stays on the host, bit-exact — and turns the accumulation stage
(:656-666) plus upload()'s aux merge (/root/reference/src/client/fedavgclient.py:158-184) into one
kernel launch over flat arenas:  every (global model, parameter) pair of the round is an *output*, every
client tensor with a non-zero coefficient in at least one output is read exactly once.
"""
import re
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .arena import MatSpec

MAX_OUT = 4
LERP, WSUM = 0, 1
SRC_PLAIN, SRC_HOLD, SRC_MERGE = 0, 1, 2


# ---- name bookkeeping (fedavgserver.py:94-115, 183-238) ---------------------------------------------
def get_name_type(name):
    if "embeddings" in name:
        return "embedding"
    if "attention" in name or "attn" in name:
        return "attn"
    if "blocks" in name:
        return "blocks"
    if "mlp" in name:
        return "mlp"
    return "task"


_num = re.compile(r"\d+")


def get_first_number(s):
    m = _num.search(s)
    return int(m.group()) if m else None


def get_name_modality(name, modalities):
    i = get_first_number(name)
    return modalities[i] if i is not None else None


def init_param_scope(names, shared_param, share_scope):
    target = {"none": None, "attn": "attn", "blocks": "blocks", "mlp": "mlp"}[shared_param]
    return {n: (share_scope if target is not None and get_name_type(n) == target else "dataset") for n in names}


@dataclass
class GlobalCtx:
    dataset: str
    modality: str
    task: str
    out_modality_scale: float
    spec: MatSpec
    arena_in: torch.Tensor          # old global (flat fp32)
    arena_out: torch.Tensor         # new global (may be the same tensor)
    out_offset_of: Optional[Dict[str, int]] = None   # key -> float offset inside arena_out when it is NOT laid out
                                                     # like the spec's arena (the compact multi-rank staging buffer)


@dataclass
class ClientCtx:
    id: int
    dataset: str
    modality: str
    task: str
    size: int
    spec: MatSpec
    arena: Optional[torch.Tensor]   # None for clients held by another rank (they only enter the normalisers)


def _client_coefs(scope, pm, g: GlobalCtx, clients: List[ClientCtx], modalities, share_scope_flag, compensation,
                  fedavg, stale_id=None):
    """{client id: python float} for one (global, scope, param-modality) class — fedavgserver.py:601-653.

    `stale_id`: the value the reference's loop variable `identifier` still holds at :648 — the LAST key of its
    `updated_sizes` dict.  Called directly with an ascending dict (CreamflServer, the golden tests) that is the
    highest id; inside `update()` the dict is `dict(ChainMap(*results))`, which lists clients in REVERSED completion
    order, so with sequential clients (`--num_thread 1`) it is the lowest sampled id, and with worker threads it is
    whichever client finished first (the reference is racy there).  None = highest id."""
    sizes = {c.id: c.size for c in clients}
    byid = {c.id: c for c in clients}
    num = {}
    old_sum = sum(sizes.values())
    for c in clients:
        n = c.size
        if scope == "all":
            num[c.id] = n
        elif scope == "dataset":
            num[c.id] = n if c.dataset == g.dataset else 0
        elif scope == "task":
            num[c.id] = n if c.task == g.task else 0
        elif scope == "modality":
            if fedavg:
                num[c.id] = n if c.modality == g.modality else 0
            else:
                num[c.id] = n if (c.modality in g.modality or g.modality in c.modality) else 0
        elif scope == "modality_exact":
            if fedavg:
                continue       # the reference's fedavg=True branch has no modality_exact case
            num[c.id] = n if (c.modality == pm or pm in c.modality) else 0
        if not fedavg and c.modality != g.modality and g.out_modality_scale != 1:
            old_sum -= num[c.id]
            num[c.id] *= g.out_modality_scale
            old_sum += num[c.id]
    if compensation and not fedavg:
        if share_scope_flag == "all":
            return {k: float(v / old_sum) for k, v in num.items()}
        if share_scope_flag == "modality" or (share_scope_flag == "modality_exact" and not pm):
            comp = sum(s for i, s in sizes.items()
                       if byid[i].modality in g.modality or g.modality in byid[i].modality)
        elif share_scope_flag == "modality_exact":
            last = byid[stale_id] if stale_id is not None else clients[-1]    # stale `identifier` (:648), reproduced
            comp = sum(s for i, s in sizes.items() if byid[i].modality == pm or pm in last.modality)
        else:
            raise KeyError("--compensation needs share_scope in {all, modality, modality_exact} "
                           "(the reference leaves coefficients unset otherwise, fedavgserver.py:640-651)")
        return {k: float(v / comp) if comp != 0 else 0 for k, v in num.items()}
    tot = sum(num.values())
    return {k: float(v / tot) if tot != 0 else 0 for k, v in num.items()}


def upload_keys(spec: MatSpec, with_aux_flag: bool, modality: str):
    """Keys present in FedavgClient.upload() (fedavgclient.py:158-184) and, per key, the aux partner."""
    merged = with_aux_flag and modality != "img+txt"
    out = {}
    names = spec.aux_layer_names() if merged else ()
    for s in spec.segments:
        k = s.key
        if merged and ("aux" in k or "cross_modal_scale" in k):
            continue
        aux = None
        if merged and any(n in k for n in names) and "weight" in k:
            ak, sk = k.replace("weight", "aux_weight"), k.replace("weight", "cross_modal_scale")
            if ak in spec._by_key:
                aux = (spec.seg(ak).offset, spec.seg(sk).offset)
            else:
                raise KeyError(ak)    # the reference raises the same KeyError (aux_*_only mismatch)
        out[k] = (s.offset, s.numel, aux)
    return out


_PLAN_CACHE = {}          # structural signature -> symbolic tables (addresses as (owner, byte offset))
_PLAN_CACHE_MAX = 64


class AggregationPlan:
    """Tables for one fc_aggregate launch (see include/fedcola_b200.h).

    The tables depend on the sampled clients only through their (spec, dataset, modality, task, size) in id order,
    so they are built symbolically — every address as (owner, byte offset) — memoised on that signature, and turned
    into device addresses with a few vectorised numpy operations per round."""

    def __init__(self, globals_: List[GlobalCtx], clients: List[ClientCtx], param_scope: Dict[str, str],
                 args_modalities, share_scope_flag, compensation, with_aux, mode=LERP, fedavg=False,
                 include_global_term=True, stale_id=None):
        clients = sorted(clients, key=lambda c: c.id)
        self.mode = mode
        if mode == LERP and any(c.arena is None for c in clients):
            raise ValueError("sequential-lerp aggregation needs every sampled client's arena on this GPU; "
                             "use mode=WSUM for clients sharded across ranks")
        # (the tables depend on client ids only through their order: key the stale identifier by its position)
        stale_pos = None if stale_id is None else [c.id for c in clients].index(stale_id)
        sig = (mode, bool(fedavg), bool(include_global_term), stale_pos, tuple(args_modalities), share_scope_flag,
               bool(compensation), bool(with_aux), id(param_scope), len(param_scope),
               tuple((g.spec.signature, g.dataset, g.modality, g.task, g.out_modality_scale, g.out_offset_of is not None)
                     for g in globals_),
               tuple((c.spec.signature, c.dataset, c.modality, c.task, c.size, c.arena is None) for c in clients))
        sym = _PLAN_CACHE.get(sig)
        if sym is None:
            sym = self._build_symbolic(globals_, clients, param_scope, args_modalities, share_scope_flag, compensation,
                                       with_aux, mode, fedavg, include_global_term, stale_id)
            sym["_pin"] = param_scope            # the key holds its id(): keep it alive
            if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
                _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
            _PLAN_CACHE[sig] = sym
        self.job_names = sym["job_names"]
        self.algorithmic_bytes = sym["algorithmic_bytes"]
        self.n_jobs, self.n_tiles = sym["n_jobs"], sym["n_tiles"]
        # ---- addresses of this round's arenas ----
        gin = np.asarray([g.arena_in.data_ptr() for g in globals_] + [0], dtype=np.uint64)
        gout = np.asarray([g.arena_out.data_ptr() for g in globals_] + [0], dtype=np.uint64)
        cbase = np.asarray([c.arena.data_ptr() if c.arena is not None else 0 for c in clients] + [0], dtype=np.uint64)
        self.host = dict(
            job_tile_start=sym["job_tile_start"], job_numel=sym["job_numel"], job_nout=sym["job_nout"],
            job_gin=gin[sym["job_g"]] + sym["job_g_off"], job_gout=gout[sym["job_g"]] + sym["job_g_off_out"],
            job_gscale=sym["job_gscale"], job_src_start=sym["job_src_start"],
            src_ptr=cbase[sym["src_owner"]] + sym["src_off"], src_flag=sym["src_flag"],
            scale_ptr=cbase[sym["scale_owner"]] + sym["scale_off"], coef=sym["coef"],
        )
        self.dev = None
        self._keep = (globals_, clients)

    @staticmethod
    def _build_symbolic(globals_, clients, param_scope, args_modalities, share_scope_flag, compensation, with_aux,
                        mode, fedavg, include_global_term, stale_id=None):
        tile = int(_lib.lib().fc_aggregate_tile_floats())
        coef_cache = {}
        ukeys = {c.id: upload_keys(c.spec, with_aux, c.modality) for c in clients}
        pos = {c.id: i for i, c in enumerate(clients)}
        NONE_G, NONE_C = len(globals_), len(clients)      # index of the all-zero sentinel address

        # union of output parameter names, in first-seen order
        names, outs = [], {}
        for gi, g in enumerate(globals_):
            for k in g.spec.required_keys():
                kk = (k, g.spec.seg(k).numel)     # same name, different shape (vocab / classes) = separate jobs
                if kk not in outs:
                    outs[kk] = []
                    names.append(kk)
                outs[kk].append(gi)

        job_numel, job_nout, job_g, job_g_off, job_g_off_out, job_gscale = [], [], [], [], [], []
        job_src_start, src_owner, src_off, src_flag, scale_owner, scale_off, coef = [0], [], [], [], [], [], []
        job_names = []
        algorithmic_bytes = 0
        for name, numel in names:
            glist = outs[(name, numel)]
            for g0 in range(0, len(glist), MAX_OUT):
                gs = glist[g0:g0 + MAX_OUT]
                cs = []
                for gi in gs:
                    g = globals_[gi]
                    scope = param_scope[name]
                    pm = None if fedavg else get_name_modality(name, args_modalities)
                    ck = (gi, scope, pm if (scope == "modality_exact" or
                                            (compensation and share_scope_flag == "modality_exact")) else None)
                    if ck not in coef_cache:
                        coef_cache[ck] = _client_coefs(scope, pm, g, clients, args_modalities, share_scope_flag,
                                                       compensation, fedavg, stale_id)
                    cs.append(coef_cache[ck])
                rows = []
                for c in clients:
                    if name not in ukeys[c.id]:
                        continue
                    off, n, aux = ukeys[c.id][name]
                    if n != numel:
                        continue      # e.g. word embeddings of a different vocabulary: load_state_dict would raise
                    cvals = [np.float32(cc[c.id]) for cc in cs]
                    if not any(v != 0 for v in cvals):
                        continue
                    rows.append((c, off, aux, cvals))
                if mode == WSUM and rows:
                    # closed form of the sequential lerp, fp64 on the host:
                    #   f = g*prod(1-c_k) + sum_k c_k*prod_{j>k}(1-c_j)*l_k
                    C = np.asarray([[float(v) for v in r[3]] for r in rows], dtype=np.float64)   # [R, nout]
                    one_minus = 1.0 - C
                    suffix = np.ones_like(C)
                    suffix[:-1] = np.cumprod(one_minus[::-1], axis=0)[::-1][1:]
                    W = C * suffix
                    wg = np.prod(one_minus, axis=0)
                    rows = [(c, off, aux, [np.float32(x) for x in W[r]]) for r, (c, off, aux, _) in enumerate(rows)]
                    gsc = [np.float32(x) if include_global_term else np.float32(0) for x in wg]
                elif mode == WSUM:
                    gsc = [np.float32(1.0 if include_global_term else 0.0)] * len(gs)
                else:
                    gsc = [np.float32(0)] * len(gs)
                rows = [r for r in rows if r[0].arena is not None]     # remote clients only shape the weights
                job_names.append(name)
                job_numel.append(numel)
                job_nout.append(len(gs))
                for o in range(MAX_OUT):
                    if o < len(gs):
                        go = globals_[gs[o]]
                        job_g.append(gs[o]), job_g_off.append(go.spec.seg(name).offset * 4)
                        job_g_off_out.append((go.out_offset_of[name] if go.out_offset_of is not None
                                              else go.spec.seg(name).offset) * 4)
                        job_gscale.append(gsc[o])
                    else:
                        job_g.append(NONE_G), job_g_off.append(0), job_g_off_out.append(0)
                        job_gscale.append(np.float32(0))
                for c, off, aux, cvals in rows:
                    cpad = cvals + [np.float32(0)] * (MAX_OUT - len(cvals))
                    if aux:      # upload() hands the server W + A*s: a HOLD entry (W) then a MERGE entry (A, s)
                        src_owner.append(pos[c.id]), src_off.append(off * 4), src_flag.append(SRC_HOLD)
                        scale_owner.append(NONE_C), scale_off.append(0)
                        coef.extend([np.float32(0)] * MAX_OUT)
                        src_owner.append(pos[c.id]), src_off.append(aux[0] * 4), src_flag.append(SRC_MERGE)
                        scale_owner.append(pos[c.id]), scale_off.append(aux[1] * 4)
                        coef.extend(cpad)
                    else:
                        src_owner.append(pos[c.id]), src_off.append(off * 4), src_flag.append(SRC_PLAIN)
                        scale_owner.append(NONE_C), scale_off.append(0)
                        coef.extend(cpad)
                    algorithmic_bytes += 4 * numel * (2 if aux else 1)
                job_src_start.append(len(src_owner))
                algorithmic_bytes += 4 * numel * 2 * len(gs)
        tiles = [(n + tile - 1) // tile for n in job_numel]
        return dict(
            job_names=job_names, algorithmic_bytes=algorithmic_bytes, n_jobs=len(job_numel), n_tiles=int(sum(tiles)),
            job_tile_start=np.concatenate([[0], np.cumsum(tiles)]).astype(np.int32),
            job_numel=np.asarray(job_numel, dtype=np.int64),
            job_nout=np.asarray(job_nout, dtype=np.int32),
            job_g=np.asarray(job_g, dtype=np.int64), job_g_off=np.asarray(job_g_off, dtype=np.uint64),
            job_g_off_out=np.asarray(job_g_off_out, dtype=np.uint64),
            job_gscale=np.asarray(job_gscale, dtype=np.float32),
            job_src_start=np.asarray(job_src_start, dtype=np.int32),
            src_owner=np.asarray(src_owner, dtype=np.int64), src_off=np.asarray(src_off, dtype=np.uint64),
            src_flag=np.asarray(src_flag, dtype=np.int32),
            scale_owner=np.asarray(scale_owner, dtype=np.int64), scale_off=np.asarray(scale_off, dtype=np.uint64),
            coef=np.asarray(coef, dtype=np.float32).reshape(-1),
        )

    def to_device(self, device):
        """Pack all tables into ONE pinned host buffer and copy once."""
        order = ["job_tile_start", "job_numel", "job_nout", "job_gin", "job_gout", "job_gscale", "job_src_start",
                 "src_ptr", "src_flag", "scale_ptr", "coef"]
        offs, total = {}, 0
        for k in order:
            a = self.host[k]
            total = (total + 15) // 16 * 16
            offs[k] = total
            total += max(a.nbytes, 16)
        buf = np.zeros(total, dtype=np.uint8)
        for k in order:
            a = self.host[k]
            buf[offs[k]:offs[k] + a.nbytes] = a.view(np.uint8).reshape(-1)
        self._dev_buf = torch.from_numpy(buf).to(device)      # ~100 KB: a pageable copy beats a cudaHostAlloc
        base = self._dev_buf.data_ptr()
        self.dev = {k: base + offs[k] for k in order}
        self.device = torch.device(device)
        return self

    def launch(self, grid_ctas=0):
        if self.dev is None:
            raise RuntimeError("AggregationPlan.to_device() must be called before launch()")
        L, d, vp = _lib.lib(), self.dev, _lib.c_vp
        rc = L.fc_aggregate(self.mode, self.n_jobs, self.n_tiles, vp(d["job_tile_start"]), vp(d["job_numel"]),
                            vp(d["job_nout"]), vp(d["job_gin"]), vp(d["job_gout"]), vp(d["job_gscale"]),
                            vp(d["job_src_start"]), vp(d["src_ptr"]), vp(d["src_flag"]), vp(d["scale_ptr"]),
                            vp(d["coef"]), int(grid_ctas), self.device.index or 0,
                            _lib.stream_ptr(self.device))
        _lib.check(rc, "fc_aggregate")


def shard_owner(position, world_size):
    """Rank that trains the client at `position` of the sorted sampled-id list — the reference's
    `cuda:(i % ngpu)` placement rule (fedavgserver.py:310-311), one process per GPU."""
    return position % world_size


def place_clients(costs, n_slots, rule="reference"):
    """Slot (rank, or GPU of a single process) for every position of the sorted sampled-id list.

    rule='reference': position % n_slots — the reference's only rule (fedavgserver.py:310-311).
    rule='balanced' : longest-processing-time greedy on `costs` (n_k * E * train FLOPs per sample): an
                      img+txt ViT-S sample costs 36.0 GF against 8.4 GF for a text sample, so position % n leaves
                      ranks up to ~1.8x apart (6 img + 6 txt + 4 pair clients over 8 ranks).  Sampling and ids are
                      unaffected; only where a client trains changes.  Deterministic: ties break on position."""
    n = len(costs)
    if rule == "reference" or n_slots <= 1:
        return [shard_owner(i, max(n_slots, 1)) for i in range(n)]
    if rule != "balanced":
        raise ValueError(f"unknown placement rule {rule!r} (reference | balanced)")
    load = [0.0] * n_slots
    count = [0] * n_slots
    out = [0] * n
    for i in sorted(range(n), key=lambda i: (-float(costs[i]), i)):
        r = min(range(n_slots), key=lambda r: (load[r], count[r], r))
        out[i] = r
        load[r] += float(costs[i])
        count[r] += 1
    return out


def train_flops_per_sample(spec: MatSpec, modality: str):
    """fwd+bwd FLOPs of one sample, 3 * [L(24 N d^2 + 4 N^2 d) + patch embed]  (SURVEY §8 table)."""
    d, L = spec.embed_dim, spec.depth

    def enc(N, img):
        return 3.0 * (L * (24.0 * N * d * d + 4.0 * N * N * d) + (2.0 * spec.num_patches * 768 * d if img else 0.0))
    f = 0.0
    if "img" in modality:
        f += enc(spec.num_patches + 1, True)
    if "txt" in modality:
        f += enc(spec.max_text_len, False)
    return f


_STAGE_CACHE = {}


def _staging_for(globals_):
    """Persistent compact staging buffer of the multi-rank aggregation: the `required_keys()` segments of every
    global model back to back (each padded to 32 floats) — what actually has to cross NVLink (ViT-S FedCola:
    403 MB instead of the 573 MB of whole arenas with their aux_weight copies).  Built once per set of global
    arenas; returns (buffer, per-global {key: float offset}, arena views, staging views)."""
    key = tuple((g.arena_in.data_ptr(), g.spec.signature) for g in globals_)
    st = _STAGE_CACHE.get(key)
    if st is None:
        offs, total = [], 0
        for g in globals_:
            o = {}
            for k in g.spec.required_keys():
                o[k] = total
                total += (g.spec.seg(k).numel + 31) // 32 * 32
            offs.append(o)
        buf = torch.zeros(max(total, 32), dtype=torch.float32, device=globals_[0].arena_in.device)
        dsts, srcs = [], []
        for g, o in zip(globals_, offs):
            for k, so in o.items():
                sg = g.spec.seg(k)
                dsts.append(g.arena_in[sg.offset:sg.offset + sg.numel])
                srcs.append(buf[so:so + sg.numel])
        _STAGE_CACHE.clear()                       # one live federation per process
        st = _STAGE_CACHE[key] = (buf, offs, dsts, srcs)
    return st


def sharded_aggregate(globals_, clients, param_scope, flags, dist, rank, execute=None):
    """Multi-GPU aggregation: every rank folds ITS clients (those with an arena) into closed-form
    partial sums written straight into a persistent compact staging buffer (rank 0 also adds the old-global term,
    the other ranks start from zero inside the kernel — no clone / memset), ONE all-reduce over that buffer (NCCL
    over NVLink on the GPU box; gloo in the CPU tests) finishes the sums, and one multi-tensor copy scatters them
    into the global arenas; aux_weight / cross_modal_scale / padding are never touched.  `execute(plan)` runs the
    plan tables (default: the CUDA kernel).  Returns the plan (for byte accounting)."""
    buf, offs, dsts, srcs = _staging_for(globals_)
    for g, o in zip(globals_, offs):
        g.arena_out, g.out_offset_of = buf, o
    plan = AggregationPlan(globals_, clients, param_scope, mode=WSUM, include_global_term=(rank == 0), **flags)
    if execute is None:
        plan.to_device(buf.device).launch()
    else:
        execute(plan)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    with torch.no_grad():
        torch._foreach_copy_(dsts, srcs)
    plan.allreduce_bytes = buf.numel() * 4
    return plan
