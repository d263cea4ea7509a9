"""Worker of tests/test_multigpu_gpu.py: one rank of a torchrun-launched federated round (NCCL).  Every rank saves the
arenas of the clients it trained (as they were when the aggregation started); rank 0 also saves the old and the new
global arenas, the sampled ids / sizes and the logged loss.  Not a test module (leading underscore)."""
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
    rank = dist.get_rank()
    from fedcola_b200.server import fedavgserver as fs
    from test_round_gpu import our_round
    snap = {}
    orig = fs.FedavgServer._aggregate_datasets

    def spy(self, datasets, ids, updated_sizes, fedavg=False):
        torch.cuda.synchronize()
        snap["clients"] = {i: self.clients[i].model.arena.cpu() for i in ids
                           if self._owner.get(i, 0) == self.rank and self.clients[i].model is not None}
        snap["old"] = {ds: self.global_models[ds].arena.cpu() for ds in datasets}
        snap["sizes"] = dict(updated_sizes)
        return orig(self, datasets, ids, updated_sizes, fedavg)

    fs.FedavgServer._aggregate_datasets = spy
    server, ids, datasets, _ = our_round(case, torch.device("cuda", local), placement=placement, num_thread=2)
    torch.cuda.synchronize()
    snap.update(ids=list(ids), owner=dict(server._owner), loss=server.results[1]["clients_updated"]["loss"]["avg"],
                new={ds: server.global_models[ds].arena.cpu() for ds in datasets},
                allreduce_bytes=server.last_aggregation["plan"].allreduce_bytes,
                arena_bytes=sum(server.global_models[ds].arena.numel() * 4 for ds in datasets))
    torch.save(snap, f"{out}.rank{rank}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
