"""`--algorithm fediot` is named by the reference's README and scripts but its modules are not shipped.
FedIoT aggregates with the FedAvg rule over `--shared_param blocks --share_scope modality_exact`, so the name the
reference would resolve (`src.algorithm.fediot.FediotOptimizer`) is the FedAvg optimizer."""
from .fedavg import FedavgOptimizer

FediotOptimizer = type("FediotOptimizer", (FedavgOptimizer,),
                       {"__module__": __name__, "__doc__": "FedAvg server rule, resolved for --algorithm fediot."})
