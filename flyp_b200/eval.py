"""Zero-shot prediction helper: src/models/eval.py:158 `pred = logits.argmax(dim=1)`."""
from __future__ import annotations

import torch

from . import ops


def zero_shot_argmax(image_features: torch.Tensor, class_features: torch.Tensor) -> torch.Tensor:
    """argmax_j <image_i, class_j> (ties -> lowest index, like torch.argmax).  The dot products come from the same
    tcgen05 path as the loss (bf16 x bf16 products are exact, fp32 accumulation) and the argmax is taken in the
    kernel's epilogue: the [n, n_classes] logits of src/models/eval.py:150-158 are never written."""
    return ops.argmax(image_features, class_features)
