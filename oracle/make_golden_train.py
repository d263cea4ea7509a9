"""TEST INFRASTRUCTURE ONLY — training golden vectors from the UNMODIFIED reference (see make_golden.py).

For every client kind (img / txt / img+txt, and the uni-modal kinds with --with_aux --aux_trained) on the
tiny fixture model (d=64, 2 blocks, 1 head; deterministic numpy weights and samples):
  * step-0 outputs (logits or unit features), loss, and sub-sampled gradients of every parameter, produced by
    the reference `ModalityAgnosticTransformer` + the reference's criterion;
  * the epoch loss and sub-sampled final parameters after `FedavgClient.update()` (SGD and AdamW) and
    `FedproxClient.update()` on 8 samples, B=4 (2 steps).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_model(kind):
    import helpers as H
    import timm
    from fedcola_b200.harness import make_args
    ds, aux = H.TRAIN_KINDS[kind]
    spec = H.train_spec(kind)
    args = make_args(shared_param="attn", share_scope="modality", vocab_size=spec.vocab_size, seq_len=H.TRAIN_SEQ,
                     dropout=0.0)
    model = timm.create_model("mome_d64_l2", pretrained=False, num_classes=list(spec.num_classes),
                              modalities=list(spec.modalities), args=args, tasks=list(spec.tasks), with_aux=aux,
                              aux_trained=True, aux_attn_only=False, aux_mlp_only=False)
    sd = {k: torch.from_numpy(v.copy()) for k, v in H.state_dict_of(spec, H.fill_arena(spec, 7)).items()}
    model.load_state_dict(sd, strict=True)
    return model, spec, args


def ref_step0(kind):
    import helpers as H
    model, spec, _ = ref_model(kind)
    ds, _ = H.TRAIN_KINDS[kind]
    a, b = H.make_samples(ds, 4, 11)
    model.train()
    m = H.DS_MODALITY[ds]
    if m == "img":
        out = model([a, None])[0]
        loss = torch.nn.CrossEntropyLoss()(out, b)
    elif m == "txt":
        out = model([None, a])[1]
        loss = torch.nn.CrossEntropyLoss()(out, b)
    else:
        outs = model([a, b], feat_out=True)
        loss = torch.nn.ContrastiveLoss()(*outs)
        out = torch.cat(outs, dim=0)
    loss.backward()
    res = {"out": out.detach().numpy(), "loss": np.float32(loss.item())}
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        res["g:" + k] = H.subsample(g.numpy())
        res["gn:" + k] = np.float32(g.norm().item())
    return res


def ref_update(kind, algorithm, optimizer, lr, mu=0.0, max_grad_norm=0.0):
    import helpers as H
    from fedcola_b200.harness import make_args
    model, spec, _ = ref_model(kind)
    ds, aux = H.TRAIN_KINDS[kind]
    m = H.DS_MODALITY[ds]
    args = make_args(algorithm=algorithm, optimizer=optimizer, lr=lr, mu=mu, max_grad_norm=max_grad_norm, B=4, E=1,
                     with_aux=aux, aux_trained=True, seq_len=H.TRAIN_SEQ, momentum=0.9 if optimizer == "SGD" else 0.0)
    if algorithm == "fedprox":
        from src.client.fedproxclient import FedproxClient as C
    else:
        from src.client.fedavgclient import FedavgClient as C
    crit = "CrossEntropyLoss" if m != "img+txt" else "ContrastiveLoss"
    c = C(args=args, training_set=H.TensorItems(ds, 8, 21), test_set=None, task=H.CLIENT_TASK[m],
          eval_metrics=["acc1"] if m != "img+txt" else ["f1"], modality=m, writer=None, criterion=crit)
    c.id, c.dataset, c.device = 0, ds, "cpu"
    c.model = model
    r = c.update()
    res = {"loss": np.float32(r[1]["loss"])}
    if m != "img+txt":
        res["acc1"] = np.float32(r[1]["metrics"]["acc1"])
    for k, v in c.model.state_dict().items():
        res["p:" + k] = H.subsample(v.numpy())
    up = c.upload()
    res["upload_keys"] = np.asarray(sorted(up.keys()))
    return res


UPDATE_RUNS = {
    # name: (algorithm, optimizer, lr, mu, max_grad_norm)
    "sgd": ("fedavg", "SGD", 0.05, 0.0, 0.0),
    "adamw": ("fedavg", "AdamW", 1e-3, 0.0, 0.0),
    "prox_sgd_clip": ("fedprox", "SGD", 0.05, 0.1, 1.0),
}


def ref_round(case):
    """One full `FedavgServer.update()` / `FedproxServer.update()` of the unmodified reference."""
    import random
    import helpers as H
    ref_shim.install()
    import src.server.fedavgserver as fs
    fs.VOCAB_SIZES.update(H.TINY_VOCAB)
    args, cds, datasets = H.round_args(case)
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    if args.algorithm == "fedprox":
        from src.server.fedproxserver import FedproxServer as S
    else:
        S = fs.FedavgServer
    server = S(args=args, writer=ref_shim.NullWriter(), server_dataset=(None, {}), client_datasets=cds,
               model_str=args.model_name)
    for i, ds in enumerate(datasets):
        spec = H.round_global_spec(case, ds)
        sd = {k: torch.from_numpy(v.copy()) for k, v in H.state_dict_of(spec, H.fill_arena(spec, 100 + i)).items()}
        server.global_models[ds].load_state_dict(sd, strict=True)
    server.round = 1
    ids = server.update()
    res = {"ids": np.asarray(ids), "loss_avg": np.float32(server.results[1]["clients_updated"]["loss"]["avg"])}
    for ds in datasets:
        for k, v in server.global_models[ds].state_dict().items():
            res[f"{ds}:{k}"] = H.subsample(v.detach().numpy(), 7)
    return res


def main():
    import helpers as H
    ref_shim.install()
    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    for kind in H.TRAIN_KINDS:
        torch.manual_seed(0)
        for k, v in ref_step0(kind).items():
            out[f"{kind}/step0/{k}"] = v
        for run, (alg, opt, lr, mu, clip) in UPDATE_RUNS.items():
            if run == "prox_sgd_clip" and kind not in ("img", "pair", "txt_aux"):
                continue
            for k, v in ref_update(kind, alg, opt, lr, mu, clip).items():
                out[f"{kind}/{run}/{k}"] = v
        print("train", kind)
    for case in H.ROUND_CASES:
        for k, v in ref_round(case).items():
            out[f"round/{case}/{k}"] = v
        print("round", case)
    np.savez_compressed(os.path.join(GOLDEN, "train_golden.npz"), **out)
    print("bytes", os.path.getsize(os.path.join(GOLDEN, "train_golden.npz")))


if __name__ == "__main__":
    main()
