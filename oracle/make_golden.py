"""TEST INFRASTRUCTURE ONLY — generates tests/golden/* from the UNMODIFIED reference (/root/reference).

Run in the build container (the reference does not exist on the GPU box):
    python oracle/make_golden.py

Fixtures
  tests/golden/agg_hashes.json   sha256 (first 16 hex) of every tensor of every new global model after
                                 FedavgServer._aggregate on the deterministic inputs of tests/helpers.py
                                 (bit-exact contract; inputs are numpy-RandomState generated, so only the
                                 hashes need to be stored)
  tests/golden/train_*.npz       logits / losses / sub-sampled gradients and post-step parameters of the
                                 reference's client step (floating-point contract)

The only deviation from stock reference behaviour: `fedavgserver.VOCAB_SIZES` (a lookup table,
fedavgserver.py:89-92) is extended with small vocabularies so the fixtures stay small.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _LenOnly(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        raise IndexError


def reference_server(case, seq_len=16):
    """A real FedavgServer of the unmodified reference, loaded with the deterministic case inputs."""
    import helpers as H
    from fedcola_b200.harness import make_args
    ref_shim.install()
    import src.server.fedavgserver as fs
    fs.VOCAB_SIZES.update(H.TINY_VOCAB)
    sp, sc, comp, aux, scales, datasets, clients = H.AGG_CASES[case]
    args = make_args(model_name="mome_d64_l2", datasets=list(datasets) + ["Coco"],
                     modalities=[H.DS_MODALITY[d] for d in datasets] + ["img+txt"], shared_param=sp, share_scope=sc,
                     compensation=comp, with_aux=aux, aux_trained=True, out_modality_scales=list(scales) + [1],
                     seq_len=seq_len, K=len(clients), Ks=[1], Cs=[1.0])
    cds = [(_LenOnly(n), None, H.CLIENT_TASK[H.DS_MODALITY[ds]], H.DS_MODALITY[ds], ds) for ds, n in clients]
    server = fs.FedavgServer(args=args, writer=ref_shim.NullWriter(), server_dataset=(None, {}), client_datasets=cds,
                             model_str=args.model_name)
    for i, ds in enumerate(datasets):
        spec = H.make_spec(ds, sp, sc, with_aux=aux, seq_len=seq_len)
        sd = {k: torch.from_numpy(v.copy()) for k, v in H.state_dict_of(spec, H.fill_arena(spec, 100 + i)).items()}
        server.global_models[ds].load_state_dict(sd, strict=True)
    for cid, (ds, n) in enumerate(clients):
        spec = H.make_spec(ds, sp, sc, with_aux=aux, seq_len=seq_len)
        c = server.clients[cid]
        c.download(server.global_models)
        sd = {k: torch.from_numpy(v.copy()) for k, v in H.state_dict_of(spec, H.fill_arena(spec, 200 + cid)).items()}
        c.model.load_state_dict(sd, strict=True)
    return server, args


def reference_aggregate(case, descending_sizes=False):
    """{dataset: {key: np.ndarray}} after the reference's _aggregate loop (fedavgserver.py:812-819).
    descending_sizes: hand `updated_sizes` over in descending-id key order, as `dict(ChainMap(*results))` produces
    it inside update() when clients complete sequentially (fedavgserver.py:578-579)."""
    import helpers as H
    ref_shim.install()
    import src.server.fedavgserver as fs
    server, args = reference_server(case)
    _, _, _, _, scales, datasets, clients = H.AGG_CASES[case]
    ids = list(range(len(clients)))
    sizes = {cid: n for cid, (_, n) in enumerate(clients)}
    if descending_sizes:
        sizes = {cid: sizes[cid] for cid in reversed(ids)}
    out = {}
    for i, ds in enumerate(server.global_models.keys()):
        server.global_model = server.global_models[ds]
        server.task = fs.DATASET_2_TASK[ds]
        server.modality = fs.DATASET_2_MODALITY[ds]
        server.dataset = ds
        server.out_modality_scale = args.out_modality_scales[i]
        server._aggregate(ids, sizes)
        out[ds] = {k: v.detach().numpy().copy() for k, v in server.global_model.state_dict().items()}
    return out


def main():
    import helpers as H
    os.makedirs(GOLDEN, exist_ok=True)
    hashes = {}
    for case in sorted(H.AGG_CASES):
        res = reference_aggregate(case)
        hashes[case] = {ds: {k: H.sha(v) for k, v in sd.items()} for ds, sd in res.items()}
        print("agg", case, sum(len(v) for v in hashes[case].values()), "tensors")
    with open(os.path.join(GOLDEN, "agg_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=0, sort_keys=True)
    try:
        from oracle import make_golden_train
        make_golden_train.main()
    except ImportError:
        pass


if __name__ == "__main__":
    main()
