"""`FediotServer`: `--algorithm fediot` is named by the reference's README/scripts but not shipped;
it is FedavgServer with `--shared_param blocks --share_scope modality_exact`."""
from .fedavgserver import FedavgServer


class FediotServer(FedavgServer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
