"""Restated `pypose.optim.optimizer` pieces the reference subclasses (0.6.8)."""
import torch
from torch import nn
from torch.optim import Optimizer


class Trivial(nn.Module):
    def forward(self, *args, **kwargs):
        out = *args, *kwargs.values()
        return out[0] if len(out) == 1 else out


class RobustModel(nn.Module):
    """Wraps a model; `forward` returns the list of residuals, `loss` = sum_i rho(|r_i|^2)."""
    def __init__(self, model, kernel=None, auto=False):
        super().__init__()
        self.model = model
        self.kernel = [Trivial()] if kernel is None else kernel

    def model_forward(self, input):
        if isinstance(input, dict):
            return self.model(**input)
        if isinstance(input, tuple):
            return self.model(*input)
        return self.model(input)

    def residual(self, output, target):
        return output if target is None else output - target

    def residuals(self, outputs, targets):
        if isinstance(outputs, tuple):
            targets = (None,) * len(outputs) if targets is None else targets
            return [self.residual(o, t) for o, t in zip(outputs, targets)]
        return [self.residual(outputs, targets)]

    def forward(self, input, target=None):
        return self.residuals(self.model_forward(input), target)

    def loss(self, input, target):
        residuals = self.residuals(self.model_forward(input), target)
        if len(self.kernel) > 1:
            vals = [k(r.square().sum(-1)).sum() for k, r in zip(self.kernel, residuals)]
        else:
            vals = [self.kernel[0](r.square().sum(-1)).sum() for r in residuals]
        return sum(vals)


class _Optimizer(Optimizer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def update_parameter(self, params, step):
        steps = step.split([p.numel() for p in params if p.requires_grad])
        [p.add_(d.view(p.shape)) for p, d in zip(params, steps) if p.requires_grad]


class LM(_Optimizer):
    def __init__(self, *a, **k):
        raise NotImplementedError("pypose.optim.LM (autograd Jacobian) is not restated; "
                                  "MAC-VO's Performant/Fast configs use LM_analytic (autodiff: false)")
