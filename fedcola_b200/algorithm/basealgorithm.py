"""Server-side optimizer interface (contract: /root/reference/src/algorithm/basealgorithm.py:5-15; dormant in the
live reference path — the round aggregates through FedavgServer._aggregate — but resolved by name)."""
from .._interface import abstract_interface

BaseOptimizer = abstract_interface(
    "BaseOptimizer",
    "Federated optimisation rule applied by the server to accumulated client updates.",
    attributes={},
    required=("step", "accumulate"),
)
