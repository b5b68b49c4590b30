"""nn.Module wrapper of revgrad (reference models/gradient_reversal/module.py:5-11)."""
import torch
from torch import nn

from .functional import revgrad


class GradientReversal(nn.Module):
    def __init__(self, alpha):
        super().__init__()
        self.alpha = torch.tensor(alpha, requires_grad=False)

    def forward(self, x):
        return revgrad(x, self.alpha)
