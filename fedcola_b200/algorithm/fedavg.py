"""`FedavgOptimizer` — the server-side optimizer name the reference resolves (`src.algorithm.fedavg`).

Upstream it is dormant: `_get_algorithm` is never called and the live aggregation is FedavgServer._aggregate's
sequential lerp, which this package runs as one CUDA launch (fedcola_b200.aggregation).  The class is kept
so that name resolution and third-party scripts keep working, with the delta-form contract it documents:

    accumulate(c, locals):  pending[name] += (server[name] - local[name]) * c[name]
    step():                 server[name]  -= pending[name]

The pending update lives in each tensor's `.grad`, which is what callers of the reference class inspect."""
import torch

from .basealgorithm import BaseOptimizer


def _is_bookkeeping(name):
    return "num_batches_tracked" in name


class FedavgOptimizer(BaseOptimizer):
    def __init__(self, params, **kwargs):
        self.params = params                      # {name: tensor}: the server's state_dict view

    def _pending(self):
        return [(p, p.grad) for p in self.params.values() if p.grad is not None]

    def zero_grad(self, set_to_none=False):
        for p, g in self._pending():
            if set_to_none:
                p.grad = None
                continue
            p.grad = g.detach()
            p.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            closure()
        pending = self._pending()
        if pending:
            torch._foreach_sub_([p.data for p, _ in pending], [g.data for _, g in pending])
        return self.params

    @torch.no_grad()
    def accumulate(self, mixing_coefficient, local_layers_iterator, check_if=_is_bookkeeping):
        for target, (name, local) in zip(self.params.values(), local_layers_iterator):
            if check_if(name) or name not in mixing_coefficient:
                continue
            weight = mixing_coefficient[name]
            contributes = weight != 0 and local is not None
            update = ((target - local) * weight).to(target.dtype) if contributes else torch.zeros_like(target)
            if target.grad is None:
                target.grad = update
            else:
                target.grad.add_(update)
