"""Restated `pypose.optim.solver` (0.6.8)."""
import torch
from torch import nn


class PINV(nn.Module):
    def __init__(self, atol=None, rtol=None, hermitian=False):
        super().__init__()
        self.atol, self.rtol, self.hermitian = atol, rtol, hermitian

    def forward(self, A, b):
        return torch.linalg.pinv(A, atol=self.atol, rtol=self.rtol, hermitian=self.hermitian) @ b


class Cholesky(nn.Module):
    def __init__(self, upper=False):
        super().__init__()
        self.upper = upper

    def forward(self, A, b):
        L, info = torch.linalg.cholesky_ex(A, upper=self.upper)
        assert not torch.any(torch.isnan(L)), "Cholesky decomposition failed."
        return b.cholesky_solve(L, upper=self.upper)
