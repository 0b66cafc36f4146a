"""`from flyp_b200.clip.loss import ClipLoss, gather_features` - same names as joliang17/FLYP clip/loss.py."""
from ..loss import ClipLoss, gather_features  # noqa: F401
