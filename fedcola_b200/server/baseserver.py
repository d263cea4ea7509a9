"""Server interface of the round driver (contract: /root/reference/src/server/baseserver.py:4-74 — `model`, `round`,
`clients` plus the ten methods `main.py` and the server itself call)."""
from .._interface import abstract_interface

BaseServer = abstract_interface(
    "BaseServer",
    "Central server of a federated run: owns the global model(s), the round counter and the client objects.",
    attributes={"round": 0, "model": None, "clients": None},
    required=("_init_model", "_get_algorithm", "_create_clients", "_sample_clients", "_request", "_aggregate",
              "_central_evaluate", "update", "evaluate", "finalize"),
)
