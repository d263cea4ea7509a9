"""`FediotOptimizer` — missing upstream; same shape as FedproxOptimizer."""
from .fedavg import FedavgOptimizer


class FediotOptimizer(FedavgOptimizer):
    def __init__(self, params, **kwargs):
        super().__init__(params=params, **kwargs)
