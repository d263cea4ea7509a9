"""Evaluation on the B200 path: `COCOEvaluator` — image<->caption retrieval recall@k."""
from .eval_coco import COCOEvaluator, recall_at_k  # noqa: F401
