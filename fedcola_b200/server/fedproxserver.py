"""`--algorithm fedprox` resolves `src.server.fedproxserver.FedproxServer`; aggregation is FedAvg's."""
from .fedavgserver import FedavgServer

FedproxServer = type("FedproxServer", (FedavgServer,), {"__module__": __name__, "__doc__": "FedAvg server under the FedProx name (the proximal term lives in FedproxClient)."})
