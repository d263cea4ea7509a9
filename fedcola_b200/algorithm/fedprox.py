"""FedProx differs from FedAvg on the client only (proximal term, fedproxclient.py); the server-side optimizer is the
FedAvg one under the name the reference resolves (`src.algorithm.fedprox.FedproxOptimizer`)."""
from .fedavg import FedavgOptimizer

FedproxOptimizer = type("FedproxOptimizer", (FedavgOptimizer,), {"__module__": __name__, "__doc__": "FedAvg server rule, resolved for --algorithm fedprox."})
