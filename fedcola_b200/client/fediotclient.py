"""`FediotClient`: the reference's scripts name `--algorithm fediot` (README.md:70, scripts/*.sh:14) but ship no
src/client/fediotclient.py.  FedIoT in this code base is FedAvg local training with
`shared_param=blocks, share_scope=modality_exact` aggregation, so the client is the FedAvg client."""
from .fedavgclient import FedavgClient


class FediotClient(FedavgClient):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
