"""Drop-in `FedproxClient` (mirror of /root/reference/src/client/fedproxclient.py:12-88): FedavgClient plus
the proximal term  mu * 0.5 * sum_tensors ||p - p_global||_2  (:64-67), evaluated by fc_sumsq / fc_prox_grad
inside the fused step against a frozen copy of the downloaded arena (:22-24)."""
from .fedavgclient import FedavgClient


class FedproxClient(FedavgClient):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def _prox(self):
        return float(self.args.mu), self.model.arena.clone()
