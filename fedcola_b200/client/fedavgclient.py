"""Drop-in `FedavgClient` (mirror of /root/reference/src/client/fedavgclient.py:15-190).

Same constructor, attributes and `download / update / upload / evaluate / __len__` contract, but `update()`
runs every batch as ONE native call (forward + loss + backward + clip + optimizer on the client's flat
arena, csrc/mat_driver.cu) on the client's CUDA device; the model never leaves HBM between download and
aggregation, and the per-step `loss.item()` host sync of the reference (:102) is replaced by device-side
accumulators read once per epoch."""
import copy
import inspect
import logging

import torch
from torch import nn

from .baseclient import BaseClient
from .. import runtime as R

logger = logging.getLogger(__name__)


class _StagedData:
    """The client's training set as whole tensors (pinned host memory or HBM), built once.
    Items follow the reference datasets: (x, y) or (img, ids, img_id, idx, idx)."""

    def __init__(self, dataset, modality, resident, device):
        n = len(dataset)
        cols = None
        for name_a, name_b in (("x", "ids" if modality == "img+txt" else "y"), ("a", "b")):
            if hasattr(dataset, name_a) and hasattr(dataset, name_b):
                cols = [getattr(dataset, name_a), getattr(dataset, name_b)]
                break
        if cols is None:      # generic path: materialise through __getitem__ once
            items = [dataset[i] for i in range(n)]
            cols = [torch.stack([torch.as_tensor(it[0]) for it in items]),
                    torch.stack([torch.as_tensor(it[1]) for it in items])]
        self.cols = []
        for c in cols:
            c = c.contiguous()
            if resident == "device":
                c = c.to(device)
            elif torch.cuda.is_available():
                c = c.pin_memory()
            self.cols.append(c)
        self.resident, self.device, self.n = resident, device, n

    def open(self):
        """Device view of the data for one update(): the HBM-resident tensors, or (host-resident) one async
        host->device copy of the client's set from pinned memory on the current stream — batches are then
        gathered on the GPU instead of being fancy-indexed by the CPU."""
        if self.resident == "device":
            return self.cols
        return [c.to(self.device, non_blocking=True) for c in self.cols]

    @staticmethod
    def batch(dev_cols, idx):
        """idx: list[int] (from the DataLoader's own batch sampler) -> contiguous device tensors."""
        contiguous = len(idx) > 0 and idx == list(range(idx[0], idx[0] + len(idx)))
        if contiguous:
            return [c[idx[0]:idx[0] + len(idx)] for c in dev_cols]
        ix = torch.as_tensor(idx, device=dev_cols[0].device)
        return [c.index_select(0, ix) for c in dev_cols]


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
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("fedcola_b200: clients train on a CUDA device; there is no CPU path "
                               f"(client.device = {self.device!r})")
        self.model.train()
        self.model.to(dev)
        if self.args.distributed or (self.args.mm_distributed and self.modality == "img+txt"):
            raise NotImplementedError("nn.DataParallel inside a client is replaced by client sharding across GPUs")
        resident = getattr(self.args, "data_resident", "host")
        cache = self.training_set.__dict__.setdefault("_fc_staged", {})     # shared by clients sharing a dataset
        if (resident, str(dev)) not in cache:
            cache[(resident, str(dev))] = _StagedData(self.training_set, self.modality, resident, dev)
        data = self._staged = cache[(resident, str(dev))]
        with torch.cuda.device(dev):
            trainer = self.trainer = self._make_trainer()        # fresh optimizer state every round (:63)
            spec = self.model.spec
            kind = {"img": R.LOSS_CE_IMG, "txt": R.LOSS_CE_TXT, "img+txt": R.LOSS_CONTRASTIVE}[self.modality]
            rng_mode = getattr(self.args, "droppath_rng", "fused")
            results = {}
            dev_cols = data.open()
            logger.info(f"[{self.task.upper()}] [{self.modality.upper()}] ...working on client {self.id}... ")
            for e in range(self.args.E):
                trainer.stats.zero_()
                num = seen = 0
                for idx in self.train_loader.batch_sampler:     # same sampler => same shuffling RNG as the reference
                    if num >= 2 and self.args.debug:
                        break
                    a, b = data.batch(dev_cols, idx)
                    dp = R.droppath_scales(spec, len(idx), dev, True, rng_mode)
                    if self.modality == "img":
                        trainer.step(a, None, b, kind, dp)
                    elif self.modality == "txt":
                        trainer.step(None, a, b, kind, dp)
                    else:
                        trainer.step(a, b, None, kind, dp)
                    num += 1
                    seen += len(idx)
                stats = trainer.stats.tolist()                   # ONE device->host read per epoch
                total = num * self.args.B if (self.args.debug and num >= 2) else len(self.training_set)
                res = {"loss": stats[2] / total, "metrics": {}}
                if self.modality != "img+txt":
                    res["metrics"] = {name: stats[1] / max(seen, 1) for name in self.eval_metrics if name == "acc1"}
                results[e + 1] = res
                logger.info(f"[Client {self.id}] loss: {res['loss']}" +
                            (f", acc1: {res['metrics'].get('acc1')}" if self.modality != "img+txt" else ""))
            del dev_cols
        return results

    @torch.inference_mode()
    def evaluate(self):
        """The reference marks this "Not used" and it cannot run there (`model(inputs, task=...)`, :118-153)."""
        if self.args.train_only:
            return {"loss": -1, "metrics": {"none": -1}}
        raise NotImplementedError("client-side evaluation is dead code in the reference (fedavgclient.py:118)")

    def download(self, models):
        self.model = copy.deepcopy(models[self.dataset])       # D2D copy of the flat arena

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
