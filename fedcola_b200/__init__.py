"""fedcola_b200 — B200-native (sm_100a) implementation of FedCola's federated-round hot path.

`install_as_src()` registers this package's drop-in modules under the names the reference resolves at run
time (`src.server.{alg}server`, `src.client.{alg}client`, `src.algorithm.{alg}`, `src.models.mome`,
`timm.create_model` for the mome_* factories), see INTEGRATION.md."""
__version__ = "0.1.0"


def install_as_src():
    import importlib
    import sys
    names = {}
    for alg in ("fedavg", "fedprox", "fediot"):
        names[f"src.server.{alg}server"] = f"fedcola_b200.server.{alg}server"
        names[f"src.client.{alg}client"] = f"fedcola_b200.client.{alg}client"
        names[f"src.algorithm.{alg}"] = f"fedcola_b200.algorithm.{alg}"
    names["src.models.mome"] = "fedcola_b200.models.mome"
    for alias, real in names.items():
        sys.modules[alias] = importlib.import_module(real)
    return sorted(names)
