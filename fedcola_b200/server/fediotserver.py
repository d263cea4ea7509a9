"""`FediotServer`: `--algorithm fediot` is named by the reference's README/scripts but not shipped;
it is FedavgServer run with `--shared_param blocks --share_scope modality_exact`."""
from .fedavgserver import FedavgServer

FediotServer = type("FediotServer", (FedavgServer,),
                    {"__module__": __name__, "__doc__": "FedAvg server under the FedIoT name (scope flags select the behaviour)."})
