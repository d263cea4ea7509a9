"""Worker of tests/test_multigpu_gpu.py: one rank of a torchrun-launched federated round (NCCL).  Rank 0 saves the
global arenas, the sampled ids and the logged loss to <out>.  Not a test module (leading underscore)."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    case, out, placement = sys.argv[1], sys.argv[2], sys.argv[3]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from test_round_gpu import our_round
    server, ids, datasets, _ = our_round(case, torch.device("cuda", local), placement=placement, num_thread=2)
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        torch.save({"ids": list(ids), "loss": server.results[1]["clients_updated"]["loss"]["avg"],
                    "arenas": {ds: server.global_models[ds].arena.cpu() for ds in datasets},
                    "allreduce_bytes": server.last_aggregation["plan"].allreduce_bytes,
                    "arena_bytes": sum(server.global_models[ds].arena.numel() * 4 for ds in datasets)}, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
