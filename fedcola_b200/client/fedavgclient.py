"""Drop-in `FedavgClient` (mirror of /root/reference/src/client/fedavgclient.py:15-190).

Same constructor, attributes and `download / update / upload / evaluate / __len__` contract, but `update()`
runs every batch as ONE native call (forward + loss + backward + clip + optimizer on the client's flat
arena, csrc/mat_driver.cu) on the client's CUDA device; the model never leaves HBM between download and
aggregation, and the per-step `loss.item()` host sync of the reference (:102) is replaced by device-side
accumulators read once per epoch."""
import copy
import threading
import inspect
import logging

import torch
from torch import nn

from .baseclient import BaseClient
from .. import runtime as R

logger = logging.getLogger(__name__)


def _as_model_input(t, is_float):
    """Inputs reach the kernels as raw pointers: images fp32, token ids / labels int64, contiguous."""
    t = torch.as_tensor(t)
    return t.to(torch.float32 if is_float else torch.int64).contiguous()


def epoch_batches(index_loader):
    """The batches of one epoch as index lists, drawn EXACTLY as the reference's `for batch in self.train_loader`
    draws them (fedavgclient.py:79): a real DataLoader iterator over range(n) consumes the global torch RNG the same
    way (the iterator's `_base_seed` draw, then the RandomSampler's seed draw), so shuffled batch order and every
    later draw from the global stream agree with the reference under the same seed."""
    return [t.tolist() for t in index_loader]


class _ClientData:
    """A client's training set as the hot loop consumes it.

    Tensor-backed sets (the dataset exposes whole tensors `x`/`y`, `x`/`ids` or `a`/`b`: the synthetic benchmark
    sets and the test fixtures) are staged once — pinned host memory, or HBM with args.data_resident='device'.
    Any other dataset goes through `__getitem__` for EVERY batch of every epoch, like the reference's DataLoader, so
    stochastic train-time transforms (RandomCrop / flips / ColorJitter, src/loaders/data.py:95-105) are re-drawn;
    `args.cache_dataset=True` opts into materialising it once."""

    def __init__(self, dataset, modality, resident, device, cache_items=False):
        self.dataset, self.modality, self.resident, self.device = dataset, modality, resident, device
        self.n = len(dataset)
        self.float_cols = (True, False) if modality != "txt" else (False, False)
        cols = None
        for name_a, name_b in (("x", "ids" if modality == "img+txt" else "y"), ("a", "b")):
            if hasattr(dataset, name_a) and hasattr(dataset, name_b):
                cols = [getattr(dataset, name_a), getattr(dataset, name_b)]
                break
        if cols is None and cache_items:
            items = [dataset[i] for i in range(self.n)]
            cols = [torch.stack([torch.as_tensor(it[0]) for it in items]),
                    torch.stack([torch.as_tensor(it[1]) for it in items])]
        self.cols = None
        if cols is not None:
            self.cols = []
            for c, fl in zip(cols, self.float_cols):
                c = _as_model_input(c, fl)
                if resident == "device":
                    c = c.to(device)
                elif torch.cuda.is_available() and not c.is_pinned():
                    c = c.pin_memory()
                self.cols.append(c)

    def _collate(self, idx):
        items = [self.dataset[i] for i in idx]
        out = []
        for j, fl in enumerate(self.float_cols):
            t = _as_model_input(torch.stack([torch.as_tensor(it[j]) for it in items]), fl)
            out.append(t.pin_memory() if torch.cuda.is_available() else t)
        return out

    def feed(self, batches):
        """Yields (a, b) device tensors for every index list of `batches`.

        HBM-resident sets: slices / one on-device gather.  Host-resident sets: the batch after the one being trained
        on is copied host->device on a side stream into the other of two device buffers (fc_h2d_rows: one
        cudaMemcpyAsync per run of consecutive rows), so the copies hide under compute and only two batches of
        inputs ever occupy HBM."""
        if self.cols is not None and self.resident == "device":
            # the epoch's row indices go up once, from pinned memory: a per-batch upload from a Python list is a
            # pageable copy, which makes the host wait for the stream to drain before every step
            flat = torch.tensor([i for idx in batches for i in idx], dtype=torch.int64)
            if torch.cuda.is_available():
                flat = flat.pin_memory()
            flat = flat.to(self.cols[0].device, non_blocking=True)
            at = 0
            for idx in batches:
                contiguous = len(idx) > 0 and idx == list(range(idx[0], idx[0] + len(idx)))
                if contiguous:
                    yield [c[idx[0]:idx[0] + len(idx)] for c in self.cols]
                else:
                    ix = flat[at:at + len(idx)]
                    yield [c.index_select(0, ix) for c in self.cols]
                at += len(idx)
            return
        if not batches:
            return
        from .. import _lib
        L = _lib.lib()
        dev = torch.device(self.device)
        compute = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        bmax = max(len(b) for b in batches)
        probe = self.cols if self.cols is not None else self._collate(batches[0][:1])
        slots = [[torch.empty((bmax,) + tuple(c.shape[1:]), dtype=c.dtype, device=dev) for c in probe] for _ in range(2)]
        # the allocator may hand out blocks whose previous users (kernels already enqueued on the compute stream) have
        # not run yet — safe for work on that stream, not for the copies of the side stream
        side.wait_stream(compute)
        free_ev, ready_ev, keep = [None, None], [None, None], [None, None]

        def issue(k):
            slot, idx = k % 2, batches[k]
            host = None if self.cols is not None else self._collate(idx)
            with torch.cuda.stream(side):
                if free_ev[slot] is not None:
                    side.wait_event(free_ev[slot])
                if host is not None:
                    for dst, src in zip(slots[slot], host):
                        dst[:len(idx)].copy_(src, non_blocking=True)
                    keep[slot] = host
                else:
                    ix = torch.as_tensor(idx, dtype=torch.int64)
                    for dst, src in zip(slots[slot], self.cols):
                        row = src[0].numel() * src.element_size() if src.dim() > 1 else src.element_size()
                        _lib.check(L.fc_h2d_rows(_lib.ptr(dst), _lib.ptr(src), _lib.c_vp(ix.data_ptr()), _lib.c_int(len(idx)),
                                                 _lib.c_ll(row), _lib.c_int(dev.index), _lib.c_vp(side.cuda_stream)),
                                   "fc_h2d_rows")
                ready_ev[slot] = side.record_event()

        issue(0)
        for k, idx in enumerate(batches):
            if k + 1 < len(batches):
                issue(k + 1)
            compute.wait_event(ready_ev[k % 2])
            yield [c[:len(idx)] for c in slots[k % 2]]
            free_ev[k % 2] = compute.record_event()        # the step that consumed this slot has been enqueued
        side.synchronize()


LOSS_KIND = {"img": R.LOSS_CE_IMG, "txt": R.LOSS_CE_TXT, "img+txt": R.LOSS_CONTRASTIVE}


def update_group(clients, ready=None):
    """Local training of a LOCKSTEP GROUP of clients: `FedavgClient.update()` (fedavgclient.py:55-116) for every client
    of the group, with batch s of all of them trained by ONE native call (runtime.group_step -> fc_client_step_group:
    every GEMM / attention / LayerNorm launch covers the whole group).  The reference trains the same clients side by
    side in ThreadPoolExecutor workers (fedavgserver.py:566-577); on a B200 one ViT-S client at B=112 does not fill the
    machine — 40-50 % of a GEMM launch is fixed cost — so clients of one architecture share their launches instead.

    The clients must sit on one device and hold the same model architecture (the server groups them by dataset);
    batch counts and sizes may differ (a client with fewer batches drops out of the later steps; a short last batch
    is trained in its own call).  `ready()` (optional) is called once the group's set-up — trainers, batch plans — is
    done and its first step is about to be enqueued: the server lets the worker threads set up one group at a time, so
    that the first group reaches the GPU after its own set-up and not after a third of everybody's (the set-up is
    interpreter work under the GIL).  Returns {client id: {epoch: {'loss', 'metrics'}}}."""
    c0 = clients[0]
    dev = torch.device(c0.device)
    E = c0.args.E
    for c in clients:
        if torch.device(c.device) != dev or c.model.spec.signature != c0.model.spec.signature:
            raise ValueError("clients of a lockstep group must share the device and the model architecture")
    results = {c.id: {} for c in clients}
    with torch.cuda.device(dev):
        for c in clients:
            c._begin_update(dev)
        # every client's batches for all its epochs, drawn client by client: the order in which sequential clients
        # (--num_thread 1) consume the global torch RNG in the reference
        plan = {}
        for c in clients:
            plan[c.id] = [epoch_batches(c._index_loader) for _ in range(E)]
            if c.args.debug:
                plan[c.id] = [b[:2] for b in plan[c.id]]
        rng_mode = getattr(c0.args, "droppath_rng", "fused")
        if ready is not None:
            ready()
        for e in range(E):
            feeds = {c.id: iter(c._staged.feed(plan[c.id][e])) for c in clients}
            seen = {c.id: 0 for c in clients}
            for c in clients:
                c.trainer.stats.zero_()
            for s in range(max(len(plan[c.id][e]) for c in clients)):
                buckets = {}                      # batch size -> clients that have a batch of that size at step s
                for c in clients:
                    if s < len(plan[c.id][e]):
                        buckets.setdefault(len(plan[c.id][e][s]), []).append(c)
                for n, members in buckets.items():
                    for k in range(0, len(members), R.MAX_GROUP):
                        part = members[k:k + R.MAX_GROUP]
                        for c in part:
                            a, b = next(feeds[c.id])
                            dp = R.droppath_scales(c.model.spec, n, dev, True, rng_mode)
                            if c.modality == "img":
                                c.trainer.prepare(a, None, b, LOSS_KIND["img"], dp)
                            elif c.modality == "txt":
                                c.trainer.prepare(None, a, b, LOSS_KIND["txt"], dp)
                            else:
                                c.trainer.prepare(a, b, None, LOSS_KIND["img+txt"], dp)
                            seen[c.id] += n
                        R.group_step([c.trainer for c in part])
            for c in clients:
                stats = c.trainer.stats.tolist()                 # ONE device->host read per client and epoch
                num = len(plan[c.id][e])
                total = num * c.args.B if (c.args.debug and num >= 2) else len(c.training_set)
                res = {"loss": stats[2] / total, "metrics": {}}
                if c.modality != "img+txt":
                    res["metrics"] = {name: stats[1] / max(seen[c.id], 1) for name in c.eval_metrics if name == "acc1"}
                results[c.id][e + 1] = res
                logger.info(f"[Client {c.id}] loss: {res['loss']}" +
                            (f", acc1: {res['metrics'].get('acc1')}" if c.modality != "img+txt" else ""))
            del feeds
    return results


_SHELL_LOCK = threading.Lock()
_SHELL_POOL_MAX = 64          # model objects kept per global model (their arenas stay allocated)


class FedavgClient(BaseClient):
    def __init__(self, args, training_set, test_set, task="cls", eval_metrics=["acc1"], modality="ct", writer=None,
                 criterion="CrossEntropyLoss"):
        super().__init__()
        self.args = args
        self.training_set = training_set
        self.test_set = test_set
        self.optim = torch.optim.__dict__[self.args.optimizer]      # name-resolved, as in the reference (:22)
        self.criterion_name = criterion
        if criterion not in ("CrossEntropyLoss", "ContrastiveLoss"):
            raise NotImplementedError(f"fedcola_b200: criterion {criterion!r} has no sm_100a kernel "
                                      "(supported: CrossEntropyLoss, ContrastiveLoss)")
        self.train_loader = self._create_dataloader(self.training_set, shuffle=not self.args.no_shuffle)
        # same batch size / shuffle flag over the sample indices: iterating it draws what iterating train_loader draws
        self._index_loader = self._create_dataloader(range(len(self.training_set)), shuffle=not self.args.no_shuffle)
        self.test_loader = self._create_dataloader(self.test_set, shuffle=False, test=True) \
            if self.test_set is not None else None
        self.task = task
        self.modality = modality
        self.eval_metrics = eval_metrics
        self.writer = writer
        self._staged = None
        self.trainer = None

    def _refine_optim_args(self, args):
        """Whatever `args` attributes match the optimizer's __init__ argument names (:34-42)."""
        required_args = inspect.getfullargspec(self.optim)[0]
        return {a: getattr(args, a) for a in required_args if hasattr(args, a)}

    def _create_dataloader(self, dataset, shuffle, test=True):
        if self.args.B == 0:
            self.args.B = len(self.training_set)
        return torch.utils.data.DataLoader(dataset=dataset, batch_size=self.args.B, shuffle=shuffle)

    # ---- hooks the FedProx / FedIoT subclasses override -----------------------------------------
    def _prox(self):
        return 0.0, None

    def _make_trainer(self):
        kw = self._refine_optim_args(self.args)
        mu, global_arena = self._prox()
        return R.ClientTrainer(self.model, optimizer=self.args.optimizer, lr=kw.get("lr", 1e-3),
                               betas=kw.get("betas", (0.9, 0.999)), eps=kw.get("eps", 1e-8),
                               weight_decay=kw.get("weight_decay", 0.0), momentum=kw.get("momentum", 0.0),
                               dampening=kw.get("dampening", 0.0), nesterov=kw.get("nesterov", False),
                               max_grad_norm=self.args.max_grad_norm, prox_mu=mu, global_arena=global_arena)

    def update(self):
        """Local training of this client alone (fedavgclient.py:55-116) — a lockstep group of one."""
        return update_group([self])[self.id]

    def _begin_update(self, dev):
        """Everything update() does before its epoch loop (:56-63): model to the device in train mode, data handle,
        a fresh optimizer (moments never persist across rounds)."""
        if dev.type != "cuda":
            raise RuntimeError("fedcola_b200: clients train on a CUDA device; there is no CPU path "
                               f"(client.device = {self.device!r})")
        self.model.train()
        self.model.to(dev)
        if self.args.distributed or (self.args.mm_distributed and self.modality == "img+txt"):
            raise NotImplementedError("nn.DataParallel inside a client is replaced by client sharding across GPUs")
        resident = getattr(self.args, "data_resident", "host")
        try:
            cache = self.training_set.__dict__.setdefault("_fc_staged", {})  # shared by clients sharing a dataset
        except AttributeError:
            cache = {}
        if (resident, str(dev)) not in cache:
            cache[(resident, str(dev))] = _ClientData(self.training_set, self.modality, resident, dev,
                                                      cache_items=getattr(self.args, "cache_dataset", False))
        self._staged = cache[(resident, str(dev))]
        self.trainer = self._make_trainer()
        logger.info(f"[{self.task.upper()}] [{self.modality.upper()}] ...working on client {self.id}... ")

    @torch.inference_mode()
    def evaluate(self):
        """The reference marks this "Not used" and it cannot run there (`model(inputs, task=...)`, :118-153)."""
        if self.args.train_only:
            return {"loss": -1, "metrics": {"none": -1}}
        raise NotImplementedError("client-side evaluation is dead code in the reference (fedavgclient.py:118)")

    def download(self, models):
        """fedavgclient.py:156 `self.model = copy.deepcopy(models[self.dataset])`.  The copy is one D2D copy of the flat
        arena; the model OBJECT (module tree + ~300 parameter views, ~15 ms of Python to build) is taken from the pool
        of objects released at the end of earlier rounds (`release_model`) when one is free on this client's GPU, so
        that a round's set-up does not leave the GPU waiting on the interpreter."""
        src = models[self.dataset]
        want = torch.device(self.device) if str(self.device).startswith("cuda") else src.device
        if want.type == "cuda" and want.index is None:
            want = torch.device("cuda", torch.cuda.current_device())
        shell = None
        with _SHELL_LOCK:
            pool = src.__dict__.setdefault("_shell_pool", [])
            for i, m in enumerate(pool):
                if m.device == want:
                    shell = pool.pop(i)
                    break
        self.model = shell.refill_from(src) if shell is not None else copy.deepcopy(src)

    def release_model(self, models):
        """Hand the model object back for reuse by a later `download` (called by the server when it drops the clients'
        models at the end of a round); the pool keeps at most one round's worth of objects per global model."""
        m, self.model = self.model, None
        src = models.get(self.dataset) if m is not None else None
        if src is None or getattr(m, "spec", None) is None or m.spec.signature != src.spec.signature:
            return
        m._runtime = None
        with _SHELL_LOCK:
            pool = src.__dict__.setdefault("_shell_pool", [])
            if len(pool) < _SHELL_POOL_MAX and all(m is not x for x in pool):
                pool.append(m)

    def upload(self):
        """state_dict with the aux branch merged (W + A*s) and aux keys removed (:158-184).  API-compat view:
        the server's fused aggregation reads the client's arena directly and does this merge on load."""
        sd = self.model.state_dict()
        if self.args.with_aux and self.modality != "img+txt":
            names = self.model.spec.aux_layer_names()
            new_sd = {k: v.clone() for k, v in sd.items()}
            for k, v in sd.items():
                if any(n in k for n in names) and "aux" not in k and "weight" in k:
                    new_sd[k] = v + new_sd[k.replace("weight", "aux_weight")] * new_sd[k.replace("weight", "cross_modal_scale")]
            for k in sd:
                if "aux" in k or "cross_modal_scale" in k:
                    new_sd.pop(k)
            return new_sd
        return sd

    def __len__(self):
        return len(self.training_set)

    def __repr__(self):
        return f"CLIENT < {self.id} >"
