"""Synthetic unit-norm pairs for the measurement / debugging tools (BASELINE.md section 3 inputs):
I = normalize(randn), T = normalize(0.5 I + 0.5 normalize(randn)).  The tools never touch oracle/."""
import torch
import torch.nn.functional as F


def synthetic_pairs(n, d, seed=0, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen)
    y = torch.randn(n, d, generator=gen)
    I = F.normalize(x, dim=-1)
    T = F.normalize(0.5 * I + 0.5 * F.normalize(y, dim=-1), dim=-1)
    return I.to(dtype), T.to(dtype)
