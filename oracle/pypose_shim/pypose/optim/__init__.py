from . import kernel, corrector, solver, strategy, scheduler, functional, optimizer
from .optimizer import LM
