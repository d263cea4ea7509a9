"""TEST INFRASTRUCTURE ONLY — one full federated round on the CPU with the oracle restatement
(fedcola_oracle.py): sample -> download -> local training -> upload/merge -> sequential-lerp aggregation ->
aux refresh -> lr decay.  Mirrors FedavgServer.update (/root/reference/src/server/fedavgserver.py:784-856).

Used by tests (pinned against tests/golden/train_golden.npz `round/*`, produced by the unmodified reference)
and by bench.py's cpu_baseline / `--impl reference` legs.  Never imported by the product."""
import copy
import random
import time

import numpy as np
import torch

from . import fedcola_oracle as O

DATASET_2_TASK = {"CIFAR100": "cls", "AG_NEWS": "cls", "MedMNIST": "cls", "MTSamples": "cls", "MedicalAbstracts": "cls",
                  "Flickr30k": "rtv", "Coco": "rtv"}
DATASET_2_MODALITY = {"CIFAR100": "img", "MedMNIST": "img", "AG_NEWS": "txt", "MTSamples": "txt",
                      "MedicalAbstracts": "txt", "Flickr30k": "img+txt", "Coco": "img+txt"}


class OracleModel:
    """{state_dict key: tensor} + the static shape info the functional forward needs."""

    def __init__(self, spec, state):
        self.spec = spec
        self.params = state         # aliases (share_scope == 'all') map two keys to the same tensor

    def clone(self):
        seen, out = {}, {}
        for k, v in self.params.items():
            if id(v) not in seen:
                seen[id(v)] = v.detach().clone()
            out[k] = seen[id(v)]
        return OracleModel(self.spec, out)

    def numpy_state(self):
        return {k: v.detach().numpy() for k, v in self.params.items()}


class OracleServer:
    def __init__(self, args, client_datasets, specs, init_states):
        """specs / init_states: {dataset: MatSpec-like} / {dataset: {key: np.ndarray}} for the global models."""
        self.args = args
        self.datasets = list(args.datasets[:-1])
        self.globals = {}
        for ds in self.datasets:
            sp = specs[ds]
            seen, st = {}, {}
            for s in sp.segments:
                root = s.alias_of or s.key
                if root not in seen:
                    seen[root] = torch.from_numpy(np.array(init_states[ds][root], dtype=np.float32, copy=True))
                st[s.key] = seen[root]
            self.globals[ds] = OracleModel(sp, st)
        names = []
        for g in self.globals.values():
            for k in g.spec.keys():
                if k not in names:
                    names.append(k)
        self.param_scope = O.init_param_scope(names, args.shared_param, args.share_scope)
        self.clients = [dict(id=i, train=c[0], task=c[2], modality=c[3], dataset=c[4]) for i, c in enumerate(client_datasets)]
        Cs = args.Cs if isinstance(args.Cs, list) else [args.Cs]
        if len(Cs) == 1:
            Cs = Cs * len(self.datasets)
        self.Cs = dict(zip(self.datasets, Cs))
        self.curr_lr = args.lr
        self.round = 0
        self.timing = {}

    def _batches(self, ds):
        loader = torch.utils.data.DataLoader(ds, batch_size=self.args.B, shuffle=not self.args.no_shuffle)
        return [b for b in loader]

    def update(self):
        a = self.args
        ids = O.sample_clients(self.datasets, self.Cs, [c["dataset"] for c in self.clients], a.equal_sampled, a.C, a.K)
        local, sizes, losses = {}, {}, {}
        t0 = time.perf_counter()
        n_samples = 0
        for i in ids:
            c = self.clients[i]
            m = self.globals[c["dataset"]].clone()           # download = deepcopy
            sp = m.spec
            rg = {s.key: s.requires_grad for s in sp.segments}
            _, res = O.client_update(m.params, rg, self._batches(c["train"]), c["modality"], sp.modalities, sp.num_heads,
                                     sp.depth, E=a.E, optimizer=a.optimizer, lr=self.curr_lr,
                                     weight_decay=a.weight_decay, momentum=a.momentum, nesterov=a.nesterov,
                                     max_grad_norm=a.max_grad_norm, mu=(a.mu if a.algorithm == "fedprox" else None),
                                     drop_path_rate=a.dropout, n_total=len(c["train"]))
            local[i], sizes[i], losses[i] = m, len(c["train"]), res[a.E]["loss"]
            n_samples += a.E * len(c["train"])
        t1 = time.perf_counter()
        meta = {i: dict(dataset=self.clients[i]["dataset"], modality=self.clients[i]["modality"],
                        task=self.clients[i]["task"]) for i in ids}
        agg_bytes = 0
        for gi, ds in enumerate(self.datasets):
            g = self.globals[ds]
            names = g.spec.required_keys()
            coefs = O.coefficients(names, self.param_scope, meta, sizes, ds, DATASET_2_MODALITY[ds], DATASET_2_TASK[ds],
                                   a.out_modality_scales[gi], a.modalities, a.share_scope, a.compensation)
            uploads = {i: O.upload_merge(local[i].numpy_state(), a.with_aux, self.clients[i]["modality"],
                                         a.aux_attn_only, a.aux_mlp_only) for i in ids}   # once per global, as upstream
            final = {k: g.params[k].numpy().copy() for k in names}
            final = O.aggregate_lerp(final, uploads, coefs, ids)
            with torch.no_grad():
                for k in names:
                    g.params[k].copy_(torch.from_numpy(final[k]))
            for k in names:
                agg_bytes += 8 * final[k].size
                for i in ids:
                    if k in uploads[i] and coefs[k][i] != 0:
                        merged = a.with_aux and self.clients[i]["modality"] != "img+txt" and \
                            (k.replace("weight", "aux_weight") in local[i].params) and k.endswith("weight")
                        agg_bytes += 4 * final[k].size * (2 if merged else 1)
        if a.with_aux:
            for ds in self.datasets:
                mod = DATASET_2_MODALITY[ds]
                if mod == "img+txt":
                    continue
                own = 0 if mod == "img" else 1
                other = [d for d in self.datasets if DATASET_2_MODALITY[d] == ("txt" if mod == "img" else "img")][0]
                g, src = self.globals[ds], self.globals[other]
                with torch.no_grad():
                    for k, sk in O.aux_refresh_map(g.spec.aux_keys(), own).items():
                        g.params[k].copy_(src.params[sk])
        t2 = time.perf_counter()
        if self.round % a.lr_decay_step == 0:
            self.curr_lr *= a.lr_decay
        self.timing = dict(train_s=t1 - t0, agg_s=t2 - t1, samples=n_samples, agg_bytes=agg_bytes)
        self.last_losses, self.last_sizes = losses, sizes
        return ids
