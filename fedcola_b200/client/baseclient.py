"""Client interface (contract: /root/reference/src/client/baseclient.py:5-50 — `id`, `model` and the local
update / evaluate / download / upload cycle the server drives)."""
from .._interface import abstract_interface

BaseClient = abstract_interface(
    "BaseClient",
    "One federated participant: private data, an identifier assigned by the server and the model it trains locally.",
    attributes={"id": None, "model": None},
    required=("update", "evaluate", "download", "upload", "__len__", "__repr__"),
)
