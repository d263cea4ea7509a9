"""Synthetic workloads + an `args` Namespace factory mirroring the reference CLI defaults.

The reference's configuration object is one argparse Namespace (`/root/reference/main.py:57-284`);
only the fields listed in SURVEY.md Appendix A.2 are read on the round hot path.  `make_args`
fills them with main.py's defaults so the same Namespace drives the reference (under the oracle
shim) and this package's drop-in servers/clients.

Dataset item shapes follow the reference datasets:
  img      (float32[3,224,224], int64 label)              torchvisionparser
  txt      (int64[seq_len], int64 label)                  torchtextparser
  img+txt  (img, ids, i//5, i, i)                         src/datasets/flickr30k.py:42
"""
from argparse import Namespace

import torch


def make_args(**overrides):
    a = dict(
        exp_name="synthetic", seed=1, server_device="cpu", dataset="synthetic",
        datasets=["CIFAR100", "AG_NEWS", "Coco"], modalities=["img", "txt", "img+txt"],
        Ks=[1, 1], Cs=[1.0], C=1.0, K=2, R=1, E=1, B=8,
        eval_type="local", eval_every=1000, eval_metrics=["acc1"], eval_batch_size=64, eval_fraction=1.0,
        algorithm="fedavg", optimizer="AdamW", lr=1e-4, lr_decay=1.0, lr_decay_step=20,
        weight_decay=0, momentum=0.0, nesterov=False, beta1=0.0, max_grad_norm=0.0, mu=0.01,
        criterion="CrossEntropyLoss", shared_param="none", share_scope="dataset", colearn_param="none",
        compensation=False, with_aux=False, aux_trained=False, aux_attn_only=False, aux_mlp_only=False,
        pretrained=False, vocab_size=30522, seq_len=64, dropout=0.0, no_shuffle=True, debug=False,
        distributed=False, mm_distributed=False, mp=False, num_thread=1, equal_sampled=True,
        warmup_modality="none", warmup_rounds=5, freeze_modality="none", freeze_rounds=5,
        out_modality_scales=[1, 1, 1, 1], fedavg_eval=False, train_only=True, test_size=0,
        result_path="./result", log_path="./log", use_tb=False, model_name="mome_d192_l4",
        mm_scale=100.0, multi_task=True,
        # --- flags that only this package reads (not in the reference CLI) ---
        precision="bf16",          # 'bf16' (tcgen05 kind::f16) | 'fp32' (validation mode)
    )
    a.update(overrides)
    return Namespace(**a)


class SyntheticImg(torch.utils.data.Dataset):
    def __init__(self, n, num_classes=100, seed=1):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randn(n, 3, 224, 224, generator=g)
        self.y = torch.randint(0, num_classes, (n,), generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class SyntheticTxt(torch.utils.data.Dataset):
    def __init__(self, n, seq_len=64, vocab=30522, num_classes=4, seed=2):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randint(0, vocab, (n, seq_len), generator=g)
        self.y = torch.randint(0, num_classes, (n,), generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class SyntheticPair(torch.utils.data.Dataset):
    """Flickr30k/COCO-shaped image–caption pairs (5 captions per image index pattern)."""

    def __init__(self, n, seq_len=64, vocab=7732, seed=3):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randn(n, 3, 224, 224, generator=g)
        self.ids = torch.randint(0, vocab, (n, seq_len), generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.ids[i], i // 5, i, i


DATASET_VOCAB = {"Flickr30k": 7732, "MedicalAbstracts": 20264}   # fedavgserver.py:89-92
DATASET_MODALITY = {"CIFAR100": "img", "MedMNIST": "img", "AG_NEWS": "txt", "MTSamples": "txt",
                    "MedicalAbstracts": "txt", "Flickr30k": "img+txt", "Coco": "img+txt"}
DATASET_CLASSES = {"CIFAR100": 100, "AG_NEWS": 4, "MedMNIST": 11, "MTSamples": 40, "MedicalAbstracts": 5}


def make_client_datasets(spec, seq_len=64, share=False):
    """spec: list of (dataset_name, n_samples, seed) -> the reference's `client_datasets` list of
    (train, test, task, modality, dataset_name) tuples (src/loaders/data.py:156,424).
    share=True: clients with the same (dataset, n) share one synthetic tensor set (benchmarks: dozens of
    clients, only the sampled ones are ever touched)."""
    out, cache = [], {}
    for name, n, seed in spec:
        if share and (name, n) in cache:
            out.append(cache[(name, n)])
            continue
        mod = DATASET_MODALITY[name]
        vocab = DATASET_VOCAB.get(name, 30522)
        if mod == "img":
            ds, task = SyntheticImg(n, DATASET_CLASSES[name], seed), "cls"
        elif mod == "txt":
            ds, task = SyntheticTxt(n, seq_len, vocab, DATASET_CLASSES[name], seed), "cls"
        else:
            ds, task = SyntheticPair(n, seq_len, vocab, seed), "img+txt"
        out.append((ds, None, task, mod, name))
        cache[(name, n)] = out[-1]
    return out
