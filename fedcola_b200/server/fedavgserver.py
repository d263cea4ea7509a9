"""Drop-in `FedavgServer` (mirror of /root/reference/src/server/fedavgserver.py:117-898).

Same constructor, attributes (`global_models`, `param_scope`, `clients`, `curr_lr`, `Cs`, `results`) and
`update() / evaluate() / finalize() / _aggregate()` contract.  What changes underneath:

  * global and client models are flat fp32 arenas that stay in HBM for the whole round (no cpu<->cuda
    ping-pong, no state_dict re-materialisation per global model);
  * `_aggregate` for ALL global models is ONE streaming kernel launch (fedcola_b200.aggregation ->
    csrc/aggregate.cu) that reproduces the reference's sequential lerp bit for bit;
  * with torch.distributed initialised (one process per GPU, NCCL), sampled clients are sharded across ranks
    (position i -> rank i % world_size, the reference's `cuda:(i % ngpu)` rule, :310-311; or cost-balanced with
    `args.placement='balanced'`), every rank reduces its own clients in closed form and one all-reduce over
    NVLink finishes the sum;
  * in a single process with several visible GPUs — the reference's own multi-GPU mode: worker threads, client
    i on `cuda:(i % ngpu)` (:310-311, 560-577) — clients train on their GPU and the aggregation kernel on the
    server GPU reads their arenas in place through NVLink peer access: the sequential lerp stays bit-exact.
    `args.client_devices` (a list of device strings) restricts the GPUs clients are placed on.

Client sampling, coefficient bookkeeping, lr decay and result logging are kept in Python, unchanged."""
import concurrent.futures
import threading
import gc
import json
import logging
import os
import random
from collections import ChainMap, defaultdict
from copy import deepcopy
from importlib import import_module

import numpy as np
import torch

from .baseserver import BaseServer
from .. import aggregation as agg
from ..models import mome

logger = logging.getLogger(__name__)

DATASET_2_TASK = {"BraTS": "seg", "MedMNIST": "cls", "CIFAR100": "cls", "AG_NEWS": "cls", "MTSamples": "cls",
                  "MedicalAbstracts": "cls", "Flickr30k": "rtv", "Coco": "rtv"}
DATASET_2_MODALITY = {"BraTS": "t1", "MedMNIST": "img", "CIFAR100": "img", "AG_NEWS": "txt", "MTSamples": "txt",
                      "MedicalAbstracts": "txt", "Flickr30k": "img+txt", "Coco": "img+txt"}
NUM_CLASS = {"CIFAR100": 100, "AG_NEWS": 4, "MedMNIST": 11, "MTSamples": 40, "MedicalAbstracts": 5, "Flickr30k": None,
             "Coco": None}
TASK_2_CRITERION = {"cls": "CrossEntropyLoss", "seg": "SegLoss", "img+txt": "ContrastiveLoss"}
VOCAB_SIZES = {"Flickr30k": 7732, "MedicalAbstracts": 20264}

get_name_type = agg.get_name_type
get_first_number = agg.get_first_number
get_name_modality = agg.get_name_modality


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class FedavgServer(BaseServer):
    def __init__(self, args, writer, server_dataset, client_datasets, model_str):
        super().__init__()
        self.args = args
        self.writer = writer
        self.round = 0
        d = _dist()
        self.rank, self.world_size = (d.get_rank(), d.get_world_size()) if d else (0, 1)
        if getattr(self.args, "mp", False):
            raise NotImplementedError("fedcola_b200: --mp (ProcessPoolExecutor clients, fedavgserver.py:560-562) is "
                                      "replaced by one process per GPU under torchrun; run without --mp")
        self.server_device = self._resolve_device(self.args.server_device)
        self._client_devices = self._resolve_client_devices()
        self._peer = {}
        if self.args.eval_type != "local":
            self._set_loaders(server_dataset)
        self.global_models = self._init_model(model_str)
        self._init_param_scope(args.shared_param, args.share_scope)
        self._set_evaluator()
        self.opt_kwargs = dict(lr=self.args.lr, momentum=self.args.beta1)
        self.curr_lr = self.args.lr
        self.clients = self._create_clients(client_datasets)
        self.results = defaultdict(dict)
        if type(args.Cs) != list or len(args.Cs) == 1:
            if type(args.Cs) == list:
                self.args.Cs = self.args.Cs * len(self.args.datasets)
            else:
                self.args.Cs = [self.args.Cs] * len(self.args.datasets)
        self.Cs = {dataset: C for dataset, C in zip(self.args.datasets, self.args.Cs)}
        self.last_aggregation = {}     # timing / byte accounting of the most recent aggregation (bench.py reads it)
        # Everything alive now (torch, datasets, clients) is long-lived: move it out of the collector's view so
        # the generational GC passes that the per-round model/trainer churn triggers stay cheap.  (The reference
        # instead pays an explicit gc.collect() per client per round, fedavgserver.py:670-675.)
        gc.collect()
        gc.freeze()

    # ---- devices --------------------------------------------------------------------------------------
    def _resolve_device(self, name):
        if not torch.cuda.is_available():
            raise RuntimeError("fedcola_b200: a CUDA device is required (no CPU fallback for the round hot path)")
        if self.world_size > 1:
            return torch.device("cuda", int(os.environ.get("LOCAL_RANK", self.rank % torch.cuda.device_count())))
        dev = torch.device(name if str(name).startswith("cuda") else "cuda:0")
        return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())

    def _resolve_client_devices(self):
        """GPUs that clients are placed on.  One process per GPU (torchrun): the rank's own GPU.  Single process:
        every visible GPU, as the reference does (`'cuda:%d' % (i % torch.cuda.device_count())`, :310-311), unless
        `args.client_devices` narrows the list."""
        if self.world_size > 1:
            return [str(self.server_device)]
        devs = getattr(self.args, "client_devices", None)
        if devs:
            out = []
            for d in devs:
                dv = torch.device(d)
                if dv.type != "cuda":
                    raise RuntimeError(f"fedcola_b200: client device {d!r} is not a CUDA device (no CPU path)")
                idx = dv.index if dv.index is not None else torch.cuda.current_device()
                if idx >= torch.cuda.device_count():
                    raise RuntimeError(f"fedcola_b200: client device {d!r} does not exist "
                                       f"({torch.cuda.device_count()} visible GPU(s))")
                out.append("cuda:%d" % idx)
            return out
        return ["cuda:%d" % i for i in range(torch.cuda.device_count())]

    def _client_device(self, slot):
        return self._client_devices[slot % len(self._client_devices)]

    def _peer_readable(self, device):
        """True when kernels on the server GPU can dereference memory of `device` (NVLink / PCIe peer access)."""
        idx = device.index
        if idx == self.server_device.index:
            return True
        if idx not in self._peer:
            from .. import _lib
            rc = _lib.lib().fc_enable_peer_access(int(self.server_device.index), int(idx))
            if rc not in (0, -3):
                _lib.check(rc, "fc_enable_peer_access")
            self._peer[idx] = rc == 0
            if rc != 0:
                logger.warning(f"no peer access {self.server_device} -> cuda:{idx}: client arenas from that GPU are "
                               "copied to the server GPU before aggregation")
        return self._peer[idx]

    # ---- construction ---------------------------------------------------------------------------------
    def _init_model(self, model_str):
        """One global model per dataset (:144-158), created by the name-resolved factory, placed in HBM."""
        self.args.datasets = self.args.datasets[:-1]
        models = {}
        for i, dataset in enumerate(self.args.datasets):
            self.args.vocab_size = VOCAB_SIZES[dataset] if dataset in VOCAB_SIZES else 30522
            kw = dict(pretrained=self.args.pretrained, args=self.args, with_aux=self.args.with_aux,
                      aux_trained=self.args.aux_trained, aux_attn_only=self.args.aux_attn_only,
                      aux_mlp_only=self.args.aux_mlp_only)
            m = DATASET_2_MODALITY[dataset]
            if m == "img":
                model = mome.create_model(model_str, num_classes=[NUM_CLASS[dataset], None],
                                          modalities=[self.args.modalities[i], None],
                                          tasks=[DATASET_2_TASK[dataset], None], **kw)
            elif m == "txt":
                model = mome.create_model(model_str, num_classes=[None, NUM_CLASS[dataset]],
                                          modalities=[None, self.args.modalities[i]],
                                          tasks=[None, DATASET_2_TASK[dataset]], **kw)
            elif m == "img+txt":
                model = mome.create_model(model_str, num_classes=[None, None], modalities=["img", "txt"],
                                          tasks=[DATASET_2_TASK[dataset], DATASET_2_TASK[dataset]], **kw)
            else:
                raise NotImplementedError(f"dataset modality {m!r}")
            models[dataset] = model.to(self.server_device)
        return models

    def _set_loaders(self, datasets):
        self.server_dataset = datasets[1]

    def _set_evaluator(self):
        """fedavgserver.py:177-181"""
        from ..metrics import COCOEvaluator
        evaluator = COCOEvaluator("matmul", n_crossfolds=5, extract_device=str(self.server_device),
                                  eval_device=str(self.server_device), verbose=False)
        evaluator.set_logger(logger)
        self.evaluator = evaluator

    def _init_param_scope(self, shared_param, share_scope):
        names = []
        for model in self.global_models.values():
            for key in model.state_dict().keys():
                if key not in names:
                    names.append(key)
        self.param_scope = agg.init_param_scope(names, shared_param, share_scope)

    def _get_algorithm(self, model, **kwargs):
        cls = import_module(f"..algorithm.{self.args.algorithm}", package=__package__).__dict__[
            f"{self.args.algorithm.title()}Optimizer"]
        return cls(params=model.state_dict(), **kwargs)

    def _create_clients(self, client_datasets):
        cls = import_module(f"..client.{self.args.algorithm}client", package=__package__).__dict__[
            f"{self.args.algorithm.title()}Client"]
        clients = []
        for identifier, datasets in enumerate(client_datasets):
            client = cls(args=self.args, training_set=datasets[0], test_set=datasets[1], task=datasets[2],
                         modality=datasets[3], eval_metrics=["acc1"] if datasets[2] == "cls" else ["f1"],
                         criterion=TASK_2_CRITERION[datasets[2]], writer=self.writer)
            client.id = identifier
            client.dataset = datasets[4]
            client.device = self._client_device(identifier)
            clients.append(client)
        logger.info(f"[{self.args.algorithm.upper()}] [Round: {str(self.round).zfill(4)}] ...created {len(clients)} clients!")
        return clients

    # ---- sampling (bit-exact: python `random`, unchanged logic, :282-312) ----------------------------------
    def _sample_clients(self, exclude=[]):
        if self.args.equal_sampled:
            sampled_client_ids = []
            for dataset in self.args.datasets:
                ids = [client.id for client in self.clients if client.dataset == dataset]
                num_sampled_clients = max(int(self.Cs[dataset] * len(ids)), 1)
                sampled_client_ids += sorted(random.sample(ids, num_sampled_clients))
            sampled_client_ids = sorted(sampled_client_ids)
        else:
            if exclude == []:
                num_sampled_clients = max(int(self.args.C * self.args.K), 1)
                sampled_client_ids = sorted(random.sample([i for i in range(self.args.K)], num_sampled_clients))
            else:
                num_unparticipated_clients = self.args.K - len(exclude)
                if num_unparticipated_clients == 0:
                    sampled_client_ids = sorted([i for i in range(self.args.K)])
                else:
                    num_sampled_clients = max(int(self.args.eval_fraction * num_unparticipated_clients), 1)
                    sampled_client_ids = sorted(random.sample(
                        [i for i in range(self.args.K) if i not in exclude], num_sampled_clients))
        if self.args.warmup_modality != "none" and self.round <= self.args.warmup_rounds:
            sampled_client_ids = [i for i in sampled_client_ids if self.clients[i].modality == self.args.warmup_modality]
        self._place(sampled_client_ids)
        return sampled_client_ids

    def _place(self, sampled_client_ids):
        """Where each sampled client trains: its rank (one process per GPU) or its GPU (single process).  The rule is
        the reference's position % n (:310-311) or, with args.placement='balanced', a cost-balanced assignment."""
        rule = getattr(self.args, "placement", "reference")
        costs = None
        if rule != "reference":
            costs = [len(self.clients[c]) * self.args.E *
                     agg.train_flops_per_sample(self.global_models[self.clients[c].dataset].spec, self.clients[c].modality)
                     for c in sampled_client_ids]
        n_slots = self.world_size if self.world_size > 1 else len(self._client_devices)
        slots = agg.place_clients(costs if costs is not None else [0] * len(sampled_client_ids), n_slots, rule)
        self._owner = {}
        for cid, slot in zip(sampled_client_ids, slots):
            if self.world_size > 1:
                self.clients[cid].device = str(self.server_device)
                self._owner[cid] = slot
            else:
                self.clients[cid].device = self._client_device(slot)
                self._owner[cid] = 0

    # ---- logging (:314-400) ----------------------------------------------------------------------------
    def _log_results(self, resulting_sizes, results, eval, participated, save_raw):
        losses, metrics, num_samples = [], defaultdict(list), []
        log_dict, averaged = defaultdict(dict), 0.0
        for identifier, result in results.items():
            r = result if eval else result[self.args.E]
            losses.append(r["loss"])
            for name, value in r["metrics"].items():
                metrics[name].append(value)
                if eval:
                    log_dict["Train/" + self.clients[identifier].modality + "_" + name] = value
                    averaged += value
            num_samples.append(resulting_sizes[identifier])
            logger.info(f"[{self.args.algorithm.upper()}] [{self.clients[identifier].dataset.upper()}] "
                        f"[Round: {str(self.round).zfill(4)}] [{'EVALUATE' if eval else 'UPDATE'}] [CLIENT] "
                        f"< {str(identifier).zfill(6)} > | loss: {r['loss']:.4f}")
        num_samples = np.array(num_samples).astype(float)
        for metric, value in metrics.items():
            log_dict["Test" if eval else "Training" + f"/{metric}_Avg."] = np.mean(value)
        log_dict["Test" if eval else "Training" + "/All_Avg."] = averaged / self.args.K
        self.writer.log(log_dict, self.round)
        la = np.array(losses).astype(float)
        weighted, std = la.dot(num_samples) / sum(num_samples), la.std()
        n10 = int(0.1 * len(la))
        top_i = np.argpartition(la, -n10)[-n10:] if len(la) > 1 else 0
        top = np.atleast_1d(la[top_i])
        top_mean = top.dot(np.atleast_1d(num_samples[top_i])) / num_samples[top_i].sum()
        bot_i = np.argpartition(la, max(1, n10 - 1))[:max(1, n10)] if len(la) > 1 else 0
        bot = np.atleast_1d(la[bot_i])
        bot_mean = bot.dot(np.atleast_1d(num_samples[bot_i])) / num_samples[bot_i].sum()
        result_dict = defaultdict(dict)
        result_dict["loss"] = {"avg": float(weighted), "std": float(std), "top10p_avg": float(top_mean),
                               "top10p_std": float(top.std()), "bottom10p_avg": float(bot_mean),
                               "bottom10p_std": float(bot.std())}
        if save_raw:
            result_dict["loss"]["raw"] = losses
        tag = f"Local {'Test' if eval else 'Training'} Loss " + eval * f"({'In' if participated else 'Out'})/"
        self.writer.log({tag + "Avg.": weighted, tag + "Std.": std}, self.round)
        return result_dict

    def _freeze_shared_params(self, client):
        for name, param in client.model.named_parameters():
            if self.param_scope[name] == "all":
                param.requires_grad = False

    def _unfreeze_params(self, client):
        for _, param in client.model.named_parameters():
            param.requires_grad = True

    # ---- client requests (:505-589) ----------------------------------------------------------------------
    def _request(self, ids, eval, participated, retain_model, save_raw):
        if eval:
            if self.args.train_only:
                return None
            raise NotImplementedError("client-side evaluation is dead code in the reference (fedavgclient.py:118)")

        def prepare(client):
            if client.model is None:
                client.download(self.global_models)
            client.args.lr = self.curr_lr
            if self.args.freeze_modality != "none" and client.modality == self.args.freeze_modality:
                hi = self.args.freeze_rounds + self.args.warmup_rounds
                if self.args.warmup_rounds < self.round <= hi:
                    self._freeze_shared_params(client)
                elif self.round > hi:
                    self._unfreeze_params(client)

        def update_group(group):
            """One worker = one lockstep group of clients (same GPU, same global model): their batches are trained
            by shared kernel launches (client/fedavgclient.py::update_group).  A group of one is the reference's
            per-client worker (:506-519)."""
            from ..client.fedavgclient import update_group as run
            dev = torch.device(group[0].device)
            stream = self._take_stream(dev)
            # groups are set up one at a time (downloads, trainers, batch plans: interpreter work that three threads
            # would only interleave under the GIL); the lock is handed on when the group's first step is enqueued
            lock = self._prep_lock
            lock.acquire()
            held = [True]

            def ready():
                if held[0]:
                    held[0] = False
                    lock.release()
            try:
                with torch.cuda.device(dev), torch.cuda.stream(stream):
                    for client in group:
                        prepare(client)
                    res = run(group, ready=ready)
                    stream.synchronize()
            finally:
                ready()
                self._stream_pool[str(dev)].put(stream)
            out = []
            for client in group:
                # only the arena has to survive until the aggregation: optimizer state, gradients, bf16 operands and
                # the activation workspace (3x the arena + GBs of activations) go back to the allocator now, so a GPU
                # can hold the arenas of dozens of sampled clients (BASELINE configs[3]: 64 ViT-B clients)
                client.trainer = None
                if client.model is not None:
                    client.model._runtime = None
                if not retain_model:
                    client.model = None
                out.append(({client.id: len(client.training_set)}, {client.id: res[client.id]}))
            return out

        self.__dict__.setdefault("_prep_lock", threading.Lock())
        torch.cuda.synchronize(self.server_device)       # the global arenas the clients copy from are final
        local = [self.clients[i] for i in ids if self._owner.get(i, 0) == self.rank]
        groups = self._lockstep_groups(local)
        results = []
        if self.args.num_thread > 1 and len(groups) > 1:
            with concurrent.futures.ThreadPoolExecutor(max_workers=self.args.num_thread) as pool:
                for fut in concurrent.futures.as_completed([pool.submit(update_group, g) for g in groups]):
                    results.extend(fut.result())
        else:
            for g in groups:
                results.extend(update_group(g))
        d = _dist()
        if d is not None:        # every rank logs the whole round: exchange the per-client epoch statistics
            results = self._exchange_results(d, ids, results)
        sizes = dict(ChainMap(*[r[0] for r in results])) if results else {}
        res = dict(ChainMap(*[r[1] for r in results])) if results else {}
        # The reference hands `dict(ChainMap(*results))` on: clients in REVERSED completion order (:578-579).  With
        # sequential clients that is descending id; worker threads make the order (and, through the stale
        # `identifier` at :648, the --compensation/modality_exact normaliser) racy there — here it is always the
        # sequential order, whatever args.num_thread is.
        sizes = {i: sizes[i] for i in reversed(list(ids))}
        res = {i: res[i] for i in reversed(list(ids))}
        self.round_results = res        # {client id: {epoch: {'loss', 'metrics'}}} of this round
        self.results[self.round]["clients_updated"] = self._log_results(sizes, res, eval=False, participated=True,
                                                                        save_raw=False)
        return sizes

    def _take_stream(self, dev):
        """A worker's CUDA stream, from a small per-device pool that lives as long as the server: torch's caching
        allocator keeps one block pool per stream, so re-using the same few streams every round lets every multi-GB
        activation workspace be served from cache (a fresh stream per worker meant fresh cudaMallocs every round)."""
        import queue
        pools = self.__dict__.setdefault("_stream_pool", {})
        q = pools.get(str(dev))
        if q is None:
            q = pools[str(dev)] = queue.Queue()
            for _ in range(max(1, int(self.args.num_thread))):
                q.put(torch.cuda.Stream(dev))
        return q.get()

    def _lockstep_groups(self, clients):
        """Partition this rank's sampled clients into lockstep groups: same GPU, same global model (dataset) — i.e. the
        same architecture and batch size — at most `args.client_group` (default 3, limit 4) per group, in id order.
        `args.client_group = 1` trains every client on its own, as the reference's workers do."""
        from .. import runtime as R
        size = int(getattr(self.args, "client_group", 3))
        size = max(1, min(size, R.MAX_GROUP))
        buckets = {}
        for c in clients:
            buckets.setdefault((c.device, c.dataset), []).append(c)
        groups = []
        for members in buckets.values():
            for k in range(0, len(members), size):
                groups.append(members[k:k + size])
        groups.sort(key=lambda g: g[0].id)
        return groups

    def _exchange_results(self, d, ids, results):
        """All ranks end with every sampled client's {epoch: {loss, metrics}}: one fixed-size all-reduce of a
        [clients, E, 3] tensor (each rank fills the rows of the clients it trained) instead of pickled objects."""
        ids = list(ids)
        row = {cid: i for i, cid in enumerate(ids)}
        E = self.args.E
        t = torch.zeros(len(ids), E, 3, dtype=torch.float64)
        for _, res in results:
            for cid, per_epoch in res.items():
                for e, r in per_epoch.items():
                    acc = r["metrics"].get("acc1")
                    t[row[cid], e - 1] = torch.tensor([r["loss"], 0.0 if acc is None else acc, 0.0 if acc is None else 1.0],
                                                      dtype=torch.float64)
        t = t.to(self.server_device)
        d.all_reduce(t, op=d.ReduceOp.SUM)
        t = t.cpu()
        out = []
        for cid in ids:
            per_epoch = {}
            for e in range(E):
                loss, acc, has = t[row[cid], e].tolist()
                per_epoch[e + 1] = {"loss": loss, "metrics": {"acc1": acc} if has > 0.5 else {}}
            out.append(({cid: len(self.clients[cid].training_set)}, {cid: per_epoch}))
        return out

    # ---- aggregation (:591-668) --------------------------------------------------------------------------
    def _ctx(self, datasets, in_place=True):
        gl = []
        for ds in datasets:
            i = list(self.global_models.keys()).index(ds)
            g = self.global_models[ds]
            gl.append(agg.GlobalCtx(ds, DATASET_2_MODALITY[ds], DATASET_2_TASK[ds], self.args.out_modality_scales[i],
                                    g.spec, g.arena, g.arena))
        return gl

    def _client_ctx(self, ids, updated_sizes):
        out = []
        for i in ids:
            c = self.clients[i]
            local = self._owner.get(i, 0) == self.rank and c.model is not None
            spec = c.model.spec if c.model is not None else self.global_models[c.dataset].spec
            out.append(agg.ClientCtx(i, c.dataset, c.modality, c.task, updated_sizes[i], spec,
                                     c.model.arena if local else None))
        return out

    def _aggregate_datasets(self, datasets, ids, updated_sizes, fedavg=False):
        assert set(updated_sizes.keys()) == set(ids)
        gl = self._ctx(datasets)
        cl = self._client_ctx(ids, updated_sizes)
        for c in cl:
            # single process, several GPUs (the reference's thread-per-client mode): the kernel on the server GPU
            # reads the remote arena in place over NVLink; without a peer path the arena is copied over first.
            # Either way every value is folded in ascending client id by one launch: bit-exact.
            if c.arena is not None and c.arena.device != self.server_device and not self._peer_readable(c.arena.device):
                c.arena = c.arena.to(self.server_device)
        flags = dict(args_modalities=self.args.modalities, share_scope_flag=self.args.share_scope,
                     compensation=self.args.compensation, with_aux=self.args.with_aux, fedavg=fedavg,
                     stale_id=list(updated_sizes.keys())[-1] if len(updated_sizes) else None)
        d = _dist()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(self.server_device):
            if d is None:
                plan = agg.AggregationPlan(gl, cl, self.param_scope, mode=agg.LERP, **flags).to_device(self.server_device)
                ev0.record()
                plan.launch()
                ev1.record()
            else:
                # closed-form partial sums per rank + one all-reduce over NVLink (not bit-exact: <= 1e-6 rel)
                ev0.record()
                plan = agg.sharded_aggregate(gl, cl, self.param_scope, flags, d, self.rank)
                ev1.record()
        self.last_aggregation = dict(events=(ev0, ev1), bytes=plan.algorithmic_bytes, jobs=plan.n_jobs,
                                     tiles=plan.n_tiles, plan=plan)

    def _aggregate(self, ids, updated_sizes, fedavg=False):
        """Reference signature: aggregates into `self.global_model` (the caller sets self.dataset etc., :812-819)."""
        self._aggregate_datasets([self.dataset], ids, updated_sizes, fedavg)

    def _empty_client_models(self):
        # the reference runs gc.collect() per client here (1.2 s of a 3.2 s CPU round); the flat
        # arenas have no reference cycles, so dropping the references frees them immediately
        for client in self.clients:
            if client.model is not None:
                client.release_model(self.global_models)     # the object is reused by a later download()
            client.model = None
            client.trainer = None

    def _refresh_aux(self):
        """aux_weight of every uni-modal global <- the other modality's global block weights (:821-845).
        The (destination, source) views are resolved once per arena pair and copied with one multi-tensor call."""
        cache = self.__dict__.setdefault("_aux_views", {})
        for dataset, model in self.global_models.items():
            modality = DATASET_2_MODALITY[dataset]
            if modality == "img+txt":
                continue
            other_mod, a, b = ("txt", "blockses.0", "blockses.1") if modality == "img" else ("img", "blockses.1", "blockses.0")
            other = [d for d in self.global_models if DATASET_2_MODALITY[d] == other_mod][0]
            src = self.global_models[other]
            key = (dataset, model.arena.data_ptr(), src.arena.data_ptr())
            if key not in cache:
                dsts, srcs = [], []
                for k in model.spec.aux_keys():
                    s = model.spec.seg(k)
                    o = src.spec.seg(k.replace("aux_", "").replace(a, b))
                    dsts.append(model.arena[s.offset:s.offset + s.numel])
                    srcs.append(src.arena[o.offset:o.offset + o.numel])
                cache[key] = (dsts, srcs)
            dsts, srcs = cache[key]
            if dsts:
                with torch.no_grad():
                    torch._foreach_copy_(dsts, srcs)

    # ---- the round (:784-856) ------------------------------------------------------------------------------
    def update(self):
        import time
        t0 = time.perf_counter()
        selected_ids = self._sample_clients()
        updated_sizes = self._request(selected_ids, eval=False, participated=True, retain_model=True, save_raw=False)
        t1 = time.perf_counter()
        if self.args.fedavg_eval:
            old = {d: m.arena.clone() for d, m in self.global_models.items()}
            self._aggregate_datasets(list(self.global_models.keys()), selected_ids, updated_sizes, fedavg=True)
            self._central_evaluate(fedavg=True)
            for d, m in self.global_models.items():
                m.arena.copy_(old[d])
        # all global models in one launch: every client tensor is read once for all the globals it feeds
        self._aggregate_datasets(list(self.global_models.keys()), selected_ids, updated_sizes)
        if self.args.with_aux:
            self._refresh_aux()
        if self.round % self.args.lr_decay_step == 0:
            self.curr_lr *= self.args.lr_decay
        torch.cuda.synchronize(self.server_device)
        self._empty_client_models()
        t2 = time.perf_counter()
        self.phase_ms = {"local_training": (t1 - t0) * 1e3, "aggregation_and_refresh": (t2 - t1) * 1e3}
        return selected_ids

    # ---- evaluation (:677-757, 858-868) --------------------------------------------------------------------
    @torch.no_grad()
    def _central_evaluate(self, fedavg=False):
        """fedavgserver.py:677-757: the global models on the server's test sets — uni-modal loss / acc1 through the native
        forward in eval mode, image<->caption retrieval recall through COCOEvaluator."""
        MM_METRICS = ("recall_1", "recall_5", "recall_10")
        suffix = "after" if not fedavg else ""
        for dataset, server_dataset in self.server_dataset.items():
            model = self.global_models[dataset]
            if DATASET_2_MODALITY[dataset] == "img+txt":
                self.evaluator.set_model(model)
                kw = getattr(self.args, "retrieval_eval_kwargs", {})      # (fold sizes of small synthetic test sets)
                result = self.evaluator.evaluate(torch.utils.data.DataLoader(dataset=server_dataset,
                                                                             batch_size=self.args.eval_batch_size, shuffle=True),
                                                 eval_batch_size=self.args.eval_batch_size, **kw)
                res = {}
                if "n_fold" in result:
                    for t in ("i2t", "t2i"):
                        for metric in MM_METRICS:
                            res[f"Result/Server {dataset} 1k_{t}_{metric.title()}"] = result["n_fold"][t][metric]
                    res[f"Test/Server {dataset} 1k_r@1sum"] = result["n_fold"]["t2i"]["recall_1"] + result["n_fold"]["i2t"]["recall_1"]
                for t in ("i2t", "t2i"):
                    for metric in MM_METRICS:
                        res[f"Result/Server {dataset} 5k_{t}_{metric.title()}"] = result[t][metric]
                res[f"Test/Server {dataset} 5k_r@1sum"] = result["t2i"]["recall_1"] + result["i2t"]["recall_1"]
                self.writer.log(res, self.round)
                self.results[self.round][f"server_evaluated_{dataset + suffix}"] = result
                continue
            model.eval()
            loss_sum = correct = n = 0
            loader = torch.utils.data.DataLoader(dataset=server_dataset, batch_size=self.args.B, shuffle=False)
            for inputs, targets in loader:
                inputs, targets = inputs.to(self.server_device), targets.to(self.server_device)
                out = model([inputs, None])[0] if DATASET_2_MODALITY[dataset] == "img" else model([None, inputs])[1]
                loss_sum += torch.nn.functional.cross_entropy(out, targets).item() * len(out)     # MetricManager.track
                correct += (out.argmax(1) == targets).sum().item()
                n += len(out)
            result = {"loss": loss_sum / len(server_dataset), "metrics": {"acc1": correct / max(n, 1)}}
            self.writer.log({f"Loss/Server {dataset + suffix} Loss": result["loss"]}, self.round)
            for name, value in result["metrics"].items():
                self.writer.log({f"Test/Server {dataset + suffix} {name.title()}": value}, self.round)
            self.results[self.round][f"server_evaluated_{dataset + suffix}"] = result

    def evaluate(self, excluded_ids):
        if self.args.eval_type != "global":
            self._request(range(self.args.K), eval=True, participated=False, retain_model=False,
                          save_raw=self.round == self.args.R)
        if self.args.eval_type != "local":
            self._central_evaluate()

    def finalize(self):
        if self.rank == 0:
            os.makedirs(self.args.result_path, exist_ok=True)
            with open(os.path.join(self.args.result_path, f"{self.args.exp_name}.json"), "w", encoding="utf8") as f:
                json.dump({k: v for k, v in self.results.items()}, f, indent=4)
            out = os.path.join(self.args.result_path, f"{self.args.exp_name}")
            os.makedirs(out, exist_ok=True)
            for dataset, model in self.global_models.items():
                torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()},
                           os.path.join(out, f"{dataset}.pt"))
        self.writer.finish()
