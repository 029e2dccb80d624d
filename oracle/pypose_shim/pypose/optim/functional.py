"""`pypose.optim.functional.modjac` is only used by the reference's optional Jacobian self-check
(`PyposeOptimizers.py:60-73`, off by default) and by `LM_autograd`; not restated."""


def modjac(*args, **kwargs):
    raise NotImplementedError("pypose.optim.functional.modjac is not restated")
