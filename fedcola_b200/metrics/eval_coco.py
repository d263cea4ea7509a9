"""Drop-in `COCOEvaluator` (mirror of /root/reference/src/metrics/eval_coco.py:95-465, eval_method='matmul').

Same `set_model / set_logger / extract_features / evaluate_recall / evaluate_n_fold / evaluate` contract and the same
score dictionary, but
  * features come from the native forward in eval mode (`model([img, ids], feat_out=True)`, no DropPath) and stay on the GPU;
  * the similarity matrix  q · gᵀ  ([25k, d] x [d, 5k] for COCO) is one fp32-accurate split-operand tcgen05 GEMM
    (ops.gemm_split) instead of 1 024-query `mm` batches;
  * the rank of a query's best positive is  #{gallery items scoring higher than its best positive}  — one masked
    reduction on the GPU — instead of a full sort plus a Python loop with `torch.where` per positive
    (eval_coco.py:331-334).  Identical whenever no two scores tie exactly.
"""
import numpy as np
import torch

from .. import ops


def recall_at_k(ranks, k):
    """eval_coco.py:39-46"""
    ranks = np.asarray(ranks)
    return 100.0 * len(np.where(ranks < k)[0]) / len(ranks)


class COCOEvaluator(object):
    def __init__(self, eval_method="matmul", n_crossfolds=-1, extract_device="cuda", eval_device="cuda", verbose=False):
        if eval_method != "matmul":
            raise NotImplementedError("fedcola_b200: only eval_method='matmul' (the one FedavgServer uses, fedavgserver.py:179)")
        self.eval_method = eval_method
        self.extract_device = extract_device
        self.eval_device = eval_device
        self.logger = None
        self.n_crossfolds = n_crossfolds

    def set_model(self, model):
        self.model = model
        self.n_embeddings = 1
        self.feat_size = model.embed_dim

    def set_criterion(self, criterion):
        self.criterion = criterion

    def set_logger(self, logger):
        self.logger = logger

    @torch.no_grad()
    def extract_features(self, dataloader):
        """eval_coco.py:134-240: one feature per distinct image (first occurrence), one per caption, captions re-ordered
        so that the captions of image i follow each other in image order."""
        dev = torch.device(self.extract_device if str(self.extract_device).startswith("cuda") else "cuda")
        self.model.eval()
        self.model.to(dev)
        ds = dataloader.dataset
        num_images, num_captions = ds.n_images, len(ds)
        iid_to_cls = getattr(ds, "iid_to_cls", None)
        image_features = torch.zeros(num_images, self.feat_size, device=dev)
        caption_features = torch.zeros(num_captions, self.feat_size, device=dev)
        image_classes, caption_classes = np.zeros(num_images), np.zeros(num_captions)
        image_ids_, caption_ids = np.zeros(num_images), np.zeros(num_captions)
        ci = cc = 0
        seen = set()
        for images, captions, image_ids, ann_ids, _ in dataloader:
            out = self.model([images.to(dev), captions.to(dev)], feat_out=True)
            fi, ft = out[0], out[1]
            ids = [int(i) for i in image_ids]
            n = len(ids)
            caption_features[cc:cc + n] = ft
            for k, image_id in enumerate(ids):
                cls = iid_to_cls.get(image_id, image_id) if iid_to_cls else image_id
                if image_id not in seen:
                    seen.add(image_id)
                    image_ids_[ci], image_classes[ci] = image_id, cls
                    image_features[ci] = fi[k]
                    ci += 1
                caption_ids[cc], caption_classes[cc] = int(ann_ids[k]), cls
                cc += 1
        if ci != num_images:
            raise RuntimeError("unexpected error, {} != {}".format(ci, num_images))
        if cc != num_captions:
            raise RuntimeError("unexpected error, {}, {}".format(cc, num_captions))
        if set(image_classes) != set(caption_classes):
            raise RuntimeError("unexpected error, I({}) != C({})".format(set(image_classes), set(caption_classes)))
        if not iid_to_cls:
            order = []
            for cls in image_classes:
                order.extend(np.where(caption_classes == cls)[0])
            order = np.array(order)
            caption_ids, caption_classes = caption_ids[order], caption_classes[order]
            caption_features = caption_features[torch.as_tensor(order, device=dev)]
        return {"image_features": image_features.unsqueeze(1), "caption_features": caption_features.unsqueeze(1),
                "image_sigmas": np.zeros((num_images, self.feat_size)), "caption_sigmas": np.zeros((num_captions, self.feat_size)),
                "image_ids": image_ids_, "caption_ids": caption_ids,
                "image_classes": torch.from_numpy(image_classes), "caption_classes": torch.from_numpy(caption_classes)}

    @torch.no_grad()
    def evaluate_recall(self, q_features, g_features, q_labels, g_labels, q_ids=None, g_ids=None, batch_size=1024):
        """eval_coco.py:290-351.  Returns recall@{1,5,10}, rsum, medr, meanr of the best-ranked positive."""
        if len(q_features) != len(q_labels):
            raise RuntimeError("length mismatch {}, {}".format(q_features.shape, q_labels.shape))
        if len(g_features) != len(g_labels):
            raise RuntimeError("length mismatch {}, {}".format(g_features.shape, g_labels.shape))
        dev = torch.device(self.eval_device if str(self.eval_device).startswith("cuda") else "cuda")
        q = q_features.reshape(len(q_labels), -1).to(dev, torch.float32).contiguous()
        g = g_features.reshape(len(g_labels), -1).to(dev, torch.float32).contiguous()
        d = q.shape[1]
        if d % 8:                                   # the GEMM wants 16-byte rows: zero-pad the feature dimension
            pad = 8 - d % 8
            q = torch.nn.functional.pad(q, (0, pad))
            g = torch.nn.functional.pad(g, (0, pad))
        ng = g.shape[0]
        ngp = (ng + 7) // 8 * 8                     # ... and an output width that is a multiple of 8
        if ngp != ng:
            g = torch.nn.functional.pad(g, (0, 0, 0, ngp - ng))
        sims = torch.empty(q.shape[0], ngp, device=dev)
        ops.gemm_split(q, g, ops.EPI_F32, sims)     # fp32-accurate q . g^T on the tensor cores
        sims = sims[:, :ng]
        ql = torch.as_tensor(np.asarray(q_labels), device=dev)
        gl = torch.as_tensor(np.asarray(g_labels), device=dev)
        pos = ql[:, None] == gl[None, :]
        best = torch.where(pos, sims, torch.full_like(sims, -float("inf"))).max(dim=1).values
        ranks = (sims > best[:, None]).sum(dim=1).cpu().numpy().astype(np.float64)
        r1, r5, r10 = recall_at_k(ranks, 1), recall_at_k(ranks, 5), recall_at_k(ranks, 10)
        return {"recall_1": r1, "recall_5": r5, "recall_10": r10, "rsum": r1 + r5 + r10,
                "medr": np.floor(np.median(ranks)) + 1, "meanr": np.mean(ranks) + 1}

    def evaluate_n_fold(self, extracted_features, n_crossfolds, n_images_per_crossfold, n_captions_per_crossfold,
                        eval_batch_size):
        """eval_coco.py:353-403"""
        f = extracted_features
        keys = ("recall_1", "recall_5", "recall_10", "rsum", "medr", "meanr")
        acc = {t: {k: [] for k in keys} for t in ("i2t", "t2i")}
        for idx in range(n_crossfolds):
            im = slice(idx * n_images_per_crossfold, (idx + 1) * n_images_per_crossfold)
            cp = slice(idx * n_captions_per_crossfold, (idx + 1) * n_captions_per_crossfold)
            s = {"i2t": self.evaluate_recall(f["image_features"][im], f["caption_features"][cp], f["image_classes"][im],
                                             f["caption_classes"][cp], batch_size=eval_batch_size),
                 "t2i": self.evaluate_recall(f["caption_features"][cp], f["image_features"][im], f["caption_classes"][cp],
                                             f["image_classes"][im], batch_size=eval_batch_size)}
            for t, sc in s.items():
                for k, v in sc.items():
                    acc[t][k].append(v)
        return {t: {k: np.mean(np.array(v)) for k, v in sc.items()} for t, sc in acc.items()}

    @torch.no_grad()
    def evaluate(self, dataloader, n_crossfolds=None, n_images_per_crossfold=1000, n_captions_per_crossfold=5000,
                 eval_batch_size=1024, key=None):
        """eval_coco.py:405-465"""
        scores = {}
        f = self.extract_features(dataloader)
        scores["mean_log_image_sigma"] = np.mean(f["image_sigmas"])
        scores["mean_log_caption_sigma"] = np.mean(f["caption_sigmas"])
        if n_crossfolds is None:
            n_crossfolds = self.n_crossfolds
        if getattr(dataloader.dataset, "iid_to_cls", None):
            n_crossfolds = -1
        if n_crossfolds > 0:
            scores["n_fold"] = self.evaluate_n_fold(f, n_crossfolds, n_images_per_crossfold, n_captions_per_crossfold,
                                                    eval_batch_size)
        scores["i2t"] = self.evaluate_recall(f["image_features"], f["caption_features"], f["image_classes"],
                                             f["caption_classes"], batch_size=eval_batch_size)
        scores["t2i"] = self.evaluate_recall(f["caption_features"], f["image_features"], f["caption_classes"],
                                             f["image_classes"], batch_size=eval_batch_size)
        for k in ("rsum", "medr", "meanr"):
            scores[k] = scores["i2t"][k] + scores["t2i"][k]
        return scores
