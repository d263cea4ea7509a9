"""`FedavgOptimizer` — the name-resolved server-side optimizer of the reference
(/root/reference/src/algorithm/fedavg.py:7-55).  Upstream it is dormant: `_get_algorithm` is never called
and the live aggregation is FedavgServer._aggregate's sequential lerp.  The class is kept
importable with the same arithmetic (delta form: grad += (server - local) * c; param -= grad) on whatever
device the state_dict lives on; the accelerated path is fedcola_b200.aggregation."""
import torch

from .basealgorithm import BaseOptimizer


class FedavgOptimizer(BaseOptimizer):
    def __init__(self, params, **kwargs):
        self.params = params

    def zero_grad(self, set_to_none=False):
        for _, p in self.params.items():
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad = p.grad.detach()
                p.grad.zero_()

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for _, p in self.params.items():
            if p.grad is not None:
                p.data.sub_(p.grad.data)
        return self.params if loss is None else self.params

    def accumulate(self, mixing_coefficient, local_layers_iterator, check_if=lambda name: "num_batches_tracked" in name):
        for server_param, (name, local) in zip(self.params.values(), local_layers_iterator):
            if check_if(name) or name not in mixing_coefficient:
                continue
            if mixing_coefficient[name] == 0 or local is None:
                delta = torch.zeros_like(server_param)
            else:
                delta = (server_param - local).mul(mixing_coefficient[name]).data.type(server_param.dtype)
            if server_param.grad is None:
                server_param.grad = delta
            else:
                server_param.grad.data.add_(delta)
