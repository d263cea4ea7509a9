"""fedcola_b200 — B200-native (sm_100a) implementation of FedCola's federated-round hot path.

`install_as_src()` registers this package's drop-in modules under the names the reference resolves at run
time (`src.server.{alg}server`, `src.client.{alg}client`, `src.algorithm.{alg}`, `src.models.mome`) and routes
`timm.create_model('mome_*', ...)` (fedavgserver.py:151-155) to this package's factories, see INTEGRATION.md."""
__version__ = "0.2.0"


def install_as_src():
    import importlib
    import sys
    import types
    names = {}
    for alg in ("fedavg", "fedprox", "fediot"):
        names[f"src.server.{alg}server"] = f"fedcola_b200.server.{alg}server"
        names[f"src.client.{alg}client"] = f"fedcola_b200.client.{alg}client"
        names[f"src.algorithm.{alg}"] = f"fedcola_b200.algorithm.{alg}"
    names["src.models.mome"] = "fedcola_b200.models.mome"
    for alias, real in names.items():
        sys.modules[alias] = importlib.import_module(real)
    # timm.create_model: the mome_* names resolve to this package; every other name goes to the real timm, if any
    from .models import mome
    try:
        import timm
        original = getattr(timm, "create_model", None)
    except ImportError:
        timm = sys.modules["timm"] = types.ModuleType("timm")
        original = None
    if getattr(original, "_fedcola_b200", False):
        original = original._original

    def create_model(model_name, pretrained=False, **kwargs):
        if model_name in mome._REGISTRY:
            return mome.create_model(model_name, pretrained=pretrained, **kwargs)
        if original is None:
            raise RuntimeError(f"Unknown model ({model_name})")
        return original(model_name, pretrained=pretrained, **kwargs)

    create_model._fedcola_b200, create_model._original = True, original
    timm.create_model = create_model
    return sorted(names) + ["timm.create_model"]
