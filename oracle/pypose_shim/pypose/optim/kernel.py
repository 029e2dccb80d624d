"""Restated `pypose.optim.kernel` (0.6.8): robust kernels act on SQUARED residual norms."""
import torch
from torch import nn


class Huber(nn.Module):
    """rho(x) = x if sqrt(x) < delta else 2*delta*sqrt(x) - delta^2."""
    def __init__(self, delta: float = 1.0):
        super().__init__()
        assert delta > 0
        self.delta = delta
        self.delta2 = delta ** 2

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        assert torch.all(input >= 0), "input has to be non-negative"
        mask = input.sqrt() < self.delta
        output = torch.zeros_like(input)
        output[mask] = input[mask]
        output[~mask] = 2 * self.delta * input[~mask].sqrt() - self.delta2
        return output
