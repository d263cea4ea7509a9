"""TEST INFRASTRUCTURE ONLY — evaluation golden values from the UNMODIFIED reference (see make_golden.py):
  * COCOEvaluator.evaluate (src/metrics/eval_coco.py) on a small synthetic retrieval set with the tiny pair model;
  * the uni-modal branch of FedavgServer._central_evaluate (loss = sum(loss_batch * len) / N, acc1) for the tiny image
    and text models.
Writes tests/golden/eval_golden.json.   python oracle/make_golden_eval.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shim  # noqa: E402
from oracle.make_golden_train import ref_model  # noqa: E402

RETRIEVAL = dict(n_images=24, caps=5, batch=16, n_crossfolds=2, n_images_per_crossfold=12, n_captions_per_crossfold=60)


def main():
    import helpers as H
    ref_shim.install()
    from src.metrics.eval_coco import COCOEvaluator
    out = {}
    model, spec, _ = ref_model("pair")
    ev = COCOEvaluator("matmul", n_crossfolds=5, extract_device="cpu", eval_device="cpu", verbose=False)
    ev.set_model(model)
    ds = H.RetrievalItems(RETRIEVAL["n_images"], RETRIEVAL["caps"])
    loader = torch.utils.data.DataLoader(ds, batch_size=RETRIEVAL["batch"], shuffle=False)
    res = ev.evaluate(loader, n_crossfolds=RETRIEVAL["n_crossfolds"], n_images_per_crossfold=RETRIEVAL["n_images_per_crossfold"],
                      n_captions_per_crossfold=RETRIEVAL["n_captions_per_crossfold"], eval_batch_size=RETRIEVAL["batch"])
    out["retrieval"] = json.loads(json.dumps(res, default=float))
    for kind, dsname in (("img", "CIFAR100"), ("txt", "AG_NEWS")):
        model, spec, _ = ref_model(kind)
        model.eval()
        a, b = H.make_samples(dsname, 40, 77)
        loss_sum = correct = 0.0
        with torch.no_grad():
            for i in range(0, 40, 8):
                o = model([a[i:i + 8], None])[0] if kind == "img" else model([None, a[i:i + 8]])[1]
                loss_sum += torch.nn.CrossEntropyLoss()(o, b[i:i + 8]).item() * len(o)
                correct += (o.argmax(1) == b[i:i + 8]).sum().item()
        out[f"central/{kind}"] = {"loss": loss_sum / 40, "acc1": correct / 40}
    with open(os.path.join(ROOT, "tests", "golden", "eval_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
