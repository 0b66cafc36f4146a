"""flyp_b200 - B200 (sm_100a) implementation of FLYP's contrastive-loss hot path (joliang17/FLYP clip/loss.py).

Public surface (mirrors the reference's operator API for this path):
    ClipLoss, gather_features          drop-in for clip/loss.py
    contrastive_cross_entropy          fused CE head of src/models/ce_ablation.py:122-123
    l2_normalize                       fused x / x.norm(dim=-1, keepdim=True) with backward
    zero_shot_argmax                   logits.argmax(dim=1) of src/models/eval.py:158 through the same tensor-core path
    project_normalize                  fused encoder tail: final projection + normalisation (clip/model.py:242-243,359,375-376)
    finetune                           the FLYP step around the operator (finetune_step, ce_ablation_step, StepLog)
"""
from ._lib import FlypError, build, load  # noqa: F401
from .loss import ClipLoss, contrastive_cross_entropy, gather_features, l2_normalize  # noqa: F401
from .eval import zero_shot_argmax  # noqa: F401
from .tail import project_normalize  # noqa: F401

__all__ = ["ClipLoss", "gather_features", "contrastive_cross_entropy", "l2_normalize", "zero_shot_argmax", "project_normalize",
           "FlypError", "build", "load"]
