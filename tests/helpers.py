"""Shared test helpers: deterministic inputs, tiny federation setups, a numpy interpreter of the
aggregation plan tables (validates the host planner on CPU — the product itself never runs on CPU)."""
import ctypes
import hashlib

import numpy as np
import torch

from fedcola_b200.arena import MatSpec
from fedcola_b200 import aggregation as agg

TINY = dict(embed_dim=64, depth=2, num_heads=1)
TINY_VOCAB = {"AG_NEWS": 512, "Flickr30k": 384, "Coco": 512}
DS_MODALITY = {"CIFAR100": "img", "AG_NEWS": "txt", "Flickr30k": "img+txt", "Coco": "img+txt"}
DS_TASK = {"CIFAR100": "cls", "AG_NEWS": "cls", "Flickr30k": "rtv", "Coco": "rtv"}
DS_CLASSES = {"CIFAR100": 100, "AG_NEWS": 4}
CLIENT_TASK = {"img": "cls", "txt": "cls", "img+txt": "img+txt"}


def make_spec(dataset, shared_param="none", share_scope="dataset", with_aux=False, aux_trained=True,
              aux_attn_only=False, aux_mlp_only=False, seq_len=16, size=TINY, vocab=None, drop_path_rate=0.0):
    m = DS_MODALITY[dataset]
    v = (vocab or TINY_VOCAB).get(dataset, 512)
    if m == "img":
        mods, ncls, tasks = ("img", None), (DS_CLASSES[dataset], None), ("cls", None)
    elif m == "txt":
        mods, ncls, tasks = (None, "txt"), (None, DS_CLASSES[dataset]), (None, "cls")
    else:
        mods, ncls, tasks = ("img", "txt"), (None, None), ("rtv", "rtv")
    return MatSpec(modalities=mods, num_classes=ncls, tasks=tasks, vocab_size=v, max_text_len=seq_len,
                   with_aux=with_aux, aux_trained=aux_trained, aux_attn_only=aux_attn_only, aux_mlp_only=aux_mlp_only,
                   share_scope=share_scope, shared_param=shared_param, drop_path_rate=drop_path_rate, **size)


def fill_arena(spec, seed, scale=0.05):
    """Deterministic (numpy RandomState) contents for every unique segment, in arena order."""
    rng = np.random.RandomState(seed)
    a = np.zeros(spec.total, dtype=np.float32)
    for s in spec.unique_segments():
        v = (rng.standard_normal(s.numel) * scale).astype(np.float32)
        if s.key.endswith("norm1.weight") or s.key.endswith("norm2.weight") or s.key == "norm.weight" \
                or s.key.endswith("LayerNorm.weight"):
            v = (1.0 + v).astype(np.float32)
        if s.key.endswith("cross_modal_scale"):
            v = np.asarray([0.01 + 0.001 * (seed % 7)], dtype=np.float32)
        a[s.offset:s.offset + s.numel] = v
    return a


def state_dict_of(spec, arena_np):
    return {s.key: arena_np[s.offset:s.offset + s.numel].reshape(s.shape) for s in spec.segments}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _view(addr, n):
    return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(addr)))


def run_plan_numpy(plan):
    """Execute AggregationPlan tables exactly like csrc/aggregate.cu does (CPU arenas only)."""
    h = plan.host
    M = agg.MAX_OUT
    for j in range(plan.n_jobs):
        n = int(h["job_numel"][j])
        nout = int(h["job_nout"][j])
        f = []
        for o in range(nout):
            g = _view(h["job_gin"][j * M + o], n)
            if plan.mode == agg.LERP:
                f.append(g.copy())
            else:
                wg = h["job_gscale"][j * M + o]
                f.append((wg * g).astype(np.float32) if wg != 0 else np.zeros(n, np.float32))
        pend = None
        for k in range(int(h["job_src_start"][j]), int(h["job_src_start"][j + 1])):
            x = _view(h["src_ptr"][k], n)
            flag = int(h["src_flag"][k])
            if flag == agg.SRC_HOLD:
                pend = x
                continue
            if flag == agg.SRC_MERGE:
                s = _view(h["scale_ptr"][k], 1)[0]
                x = (pend + (x * s).astype(np.float32)).astype(np.float32)
            for o in range(nout):
                c = h["coef"][k * M + o]
                if plan.mode == agg.LERP:
                    if c != 0:
                        t = ((x - f[o]).astype(np.float32) * c).astype(np.float32)
                        f[o] = (f[o] + t).astype(np.float32)
                else:
                    f[o] = (f[o].astype(np.float64) + np.float64(c) * x.astype(np.float64)).astype(np.float32)
        for o in range(nout):
            _view(h["job_gout"][j * M + o], n)[:] = f[o]


# ---- aggregation cases (shared by the oracle tests, the GPU parity tests and oracle/make_golden.py) ----
AGG_CASES = {
    # name: (shared_param, share_scope, compensation, with_aux, out_modality_scales, datasets(globals), clients[(dataset, n)])
    "fedavg_none_dataset": ("none", "dataset", False, False, [1, 1, 1], ["CIFAR100", "AG_NEWS"],
                            [("CIFAR100", 16), ("AG_NEWS", 24)]),
    "fedcola_attn_modality_comp_aux": ("attn", "modality", True, True, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                                       [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
    "fedcola_attn_modality_aux": ("attn", "modality", False, True, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                                  [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
    "fediot_blocks_modality_exact": ("blocks", "modality_exact", False, False, [1, 1, 1],
                                     ["CIFAR100", "AG_NEWS", "Flickr30k"],
                                     [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
    "attn_all": ("attn", "all", False, False, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                 [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
    "blocks_all_comp": ("blocks", "all", True, False, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                        [("CIFAR100", 10), ("AG_NEWS", 30), ("Flickr30k", 20)]),
    "attn_modality_scaled": ("attn", "modality", False, False, [0.5, 2, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                             [("CIFAR100", 16), ("CIFAR100", 8), ("AG_NEWS", 24), ("Flickr30k", 16), ("Flickr30k", 12)]),
    "modality_exact_comp": ("blocks", "modality_exact", True, False, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                            [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
    "no_pair_client": ("attn", "modality", True, True, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                       [("CIFAR100", 16), ("CIFAR100", 20), ("AG_NEWS", 24)]),
    "mlp_shared_is_noop": ("mlp", "modality", False, False, [1, 1, 1], ["CIFAR100", "AG_NEWS", "Flickr30k"],
                           [("CIFAR100", 16), ("AG_NEWS", 24), ("Flickr30k", 16)]),
}
ARGS_MODALITIES = {"CIFAR100": "img", "AG_NEWS": "txt", "Flickr30k": "img+txt"}


def random_agg_case(seed):
    """A seeded random aggregation configuration in AGG_CASES format (valid flag combinations only: compensation
    needs share_scope in {all, modality, modality_exact}, fedavgserver.py:640-651)."""
    import random
    r = random.Random(seed)
    sp = r.choice(["none", "attn", "blocks", "mlp"])
    sc = r.choice(["dataset", "modality", "modality_exact", "all"])
    comp = sc != "dataset" and r.random() < 0.5
    aux = sp == "attn" and sc == "modality" and r.random() < 0.85
    scales = [r.choice([1, 1, 0.5, 2]) for _ in range(3)]
    datasets = ["CIFAR100", "AG_NEWS", "Flickr30k"]
    clients = [(r.choice(datasets), r.randint(1, 40)) for _ in range(r.randint(1, 6))]
    if r.random() < 0.7:                     # most rounds see every modality
        clients += [(d, r.randint(1, 40)) for d in datasets if d not in {c[0] for c in clients}]
    clients.sort(key=lambda c: datasets.index(c[0]))          # ids are grouped by dataset in the reference
    return sp, sc, comp, aux, scales, datasets, clients


def build_agg_case(name, device="cpu", seq_len=16):
    """Returns (globals: [GlobalCtx], clients: [ClientCtx], param_scope, flags dict)."""
    sp, sc, comp, aux, scales, datasets, clients = AGG_CASES[name]
    gl, names = [], []
    for i, ds in enumerate(datasets):
        spec = make_spec(ds, sp, sc, with_aux=aux, seq_len=seq_len)
        a = torch.from_numpy(fill_arena(spec, 100 + i)).to(device)
        gl.append(agg.GlobalCtx(ds, DS_MODALITY[ds], DS_TASK[ds], scales[i], spec, a, a.clone()))
        for k in spec.keys():
            if k not in names:
                names.append(k)
    cl = []
    for cid, (ds, n) in enumerate(clients):
        spec = make_spec(ds, sp, sc, with_aux=aux, seq_len=seq_len)
        a = torch.from_numpy(fill_arena(spec, 200 + cid)).to(device)
        m = DS_MODALITY[ds]
        cl.append(agg.ClientCtx(cid, ds, m, CLIENT_TASK[m], n, spec, a))
    scope = agg.init_param_scope(names, sp, sc)
    mods = [DS_MODALITY[d] for d in datasets]
    return gl, cl, scope, dict(args_modalities=mods, share_scope_flag=sc, compensation=comp, with_aux=aux)


# ---- training cases (shared by oracle tests, GPU parity tests and oracle/make_golden_train.py) -------------
TRAIN_SEQ = 16
TRAIN_KINDS = {
    # name: (dataset, with_aux)
    "img": ("CIFAR100", False),
    "txt": ("AG_NEWS", False),
    "pair": ("Flickr30k", False),
    "img_aux": ("CIFAR100", True),
    "txt_aux": ("AG_NEWS", True),
}


def make_samples(dataset, n, seed, seq_len=TRAIN_SEQ):
    """Deterministic synthetic samples (numpy RandomState) in the reference's item layout."""
    rng = np.random.RandomState(seed)
    m = DS_MODALITY[dataset]
    vocab = TINY_VOCAB.get(dataset, 512)
    if m == "img":
        return (torch.from_numpy(rng.standard_normal((n, 3, 224, 224)).astype(np.float32)),
                torch.from_numpy(rng.randint(0, DS_CLASSES[dataset], size=n).astype(np.int64)))
    if m == "txt":
        return (torch.from_numpy(rng.randint(0, vocab, size=(n, seq_len)).astype(np.int64)),
                torch.from_numpy(rng.randint(0, DS_CLASSES[dataset], size=n).astype(np.int64)))
    return (torch.from_numpy(rng.standard_normal((n, 3, 224, 224)).astype(np.float32)),
            torch.from_numpy(rng.randint(0, vocab, size=(n, seq_len)).astype(np.int64)))


class TensorItems(torch.utils.data.Dataset):
    """Wraps make_samples output as the reference's datasets do: (x, y) or (img, ids, i//5, i, i)."""

    def __init__(self, dataset, n, seed, seq_len=TRAIN_SEQ):
        self.modality = DS_MODALITY[dataset]
        self.a, self.b = make_samples(dataset, n, seed, seq_len)

    def __len__(self):
        return self.a.shape[0]

    def __getitem__(self, i):
        if self.modality == "img+txt":
            return self.a[i], self.b[i], i // 5, i, i
        return self.a[i], self.b[i]


class RetrievalItems(torch.utils.data.Dataset):
    """COCO/Flickr-style retrieval test set: `caps` captions per image, items (img, ids, image_id, ann_id, index) and the
    attributes COCOEvaluator.extract_features reads (n_images, iid_to_cls)."""

    def __init__(self, n_images, caps=5, seed=61, seq_len=TRAIN_SEQ, vocab=384):
        rng = np.random.RandomState(seed)
        self.imgs = torch.from_numpy(rng.standard_normal((n_images, 3, 224, 224)).astype(np.float32))
        self.ids = torch.from_numpy(rng.randint(1, vocab, size=(n_images * caps, seq_len)).astype(np.int64))
        self.caps, self.n_images, self.iid_to_cls = caps, n_images, None

    def __len__(self):
        return self.ids.shape[0]

    def __getitem__(self, i):
        return self.imgs[i // self.caps], self.ids[i], 1000 + i // self.caps, 5000 + i, i


def subsample(a, step=13):
    return np.ascontiguousarray(np.asarray(a).reshape(-1)[::step])


def train_spec(kind, drop_path_rate=0.0, size=TINY):
    ds, aux = TRAIN_KINDS[kind]
    return make_spec(ds, "attn", "modality", with_aux=aux, aux_trained=True, seq_len=TRAIN_SEQ, size=size,
                     drop_path_rate=drop_path_rate)


# ---- full-round cases ----------------------------------------------------------------------------------------
ROUND_CASES = {
    # name: dict(args overrides), clients [(dataset, n, seed)]
    "fedavg": (dict(algorithm="fedavg", shared_param="none", share_scope="dataset"),
               [("CIFAR100", 8, 51), ("AG_NEWS", 12, 52)], ["CIFAR100", "AG_NEWS"]),
    "fedcola": (dict(algorithm="fedavg", shared_param="attn", share_scope="modality", compensation=True, with_aux=True,
                     aux_trained=True),
                [("CIFAR100", 8, 51), ("AG_NEWS", 12, 52), ("Flickr30k", 8, 53)], ["CIFAR100", "AG_NEWS", "Flickr30k"]),
    "fedprox": (dict(algorithm="fedprox", shared_param="none", share_scope="dataset", mu=0.1),
                [("CIFAR100", 8, 51), ("AG_NEWS", 12, 52), ("Flickr30k", 8, 53)], ["CIFAR100", "AG_NEWS", "Flickr30k"]),
    "fediot": (dict(algorithm="fedavg", shared_param="blocks", share_scope="modality_exact"),
               [("CIFAR100", 8, 51), ("CIFAR100", 4, 54), ("AG_NEWS", 12, 52), ("Flickr30k", 8, 53)],
               ["CIFAR100", "AG_NEWS", "Flickr30k"]),
    # share_scope=all: the txt encoder's attention aliases the img encoder's inside every model (mome.py:826-835)
    "attn_all": (dict(algorithm="fedavg", shared_param="attn", share_scope="all"),
                 [("CIFAR100", 8, 51), ("AG_NEWS", 12, 52), ("Flickr30k", 8, 53)], ["CIFAR100", "AG_NEWS", "Flickr30k"]),
    # the bench configuration's optimizer path: AdamW + gradient clipping, FedCola flags, two local epochs
    "fedcola_adamw_clip": (dict(algorithm="fedavg", shared_param="attn", share_scope="modality", compensation=True,
                                with_aux=True, aux_trained=True, optimizer="AdamW", lr=1e-3, max_grad_norm=1.0, E=2),
                           [("CIFAR100", 8, 51), ("AG_NEWS", 12, 52), ("Flickr30k", 8, 53)],
                           ["CIFAR100", "AG_NEWS", "Flickr30k"]),
}


def round_args(case):
    from fedcola_b200.harness import make_args
    over, clients, datasets = ROUND_CASES[case]
    kw = dict(model_name="mome_d64_l2", datasets=list(datasets) + ["Coco"],
              modalities=[DS_MODALITY[d] for d in datasets] + ["img+txt"], seq_len=TRAIN_SEQ, K=len(clients),
              Ks=[1], Cs=[1.0], B=4, E=1, optimizer="SGD", lr=0.05, momentum=0.9, out_modality_scales=[1, 1, 1, 1],
              dropout=0.0, no_shuffle=True, seed=1)
    kw.update(over)
    args = make_args(**kw)
    cds = [(TensorItems(ds, n, seed), None, CLIENT_TASK[DS_MODALITY[ds]], DS_MODALITY[ds], ds) for ds, n, seed in clients]
    return args, cds, datasets


def round_global_spec(case, ds):
    over, _, _ = ROUND_CASES[case]
    return make_spec(ds, over.get("shared_param", "none"), over.get("share_scope", "dataset"),
                     with_aux=over.get("with_aux", False), aux_trained=over.get("aux_trained", False), seq_len=TRAIN_SEQ)
