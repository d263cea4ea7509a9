"""Mirror of /root/reference/src/server/fedproxserver.py:9-11."""
from .fedavgserver import FedavgServer


class FedproxServer(FedavgServer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
