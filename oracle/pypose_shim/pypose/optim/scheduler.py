"""Restated `pypose.optim.scheduler.StopOnPlateau` (0.6.8)."""


class _Scheduler(object):
    def __init__(self, optimizer, steps, verbose=False):
        self.optimizer, self.verbose = optimizer, verbose
        self.max_steps, self.steps = steps, 0
        self._continual = True

    def continual(self):
        return self._continual


class StopOnPlateau(_Scheduler):
    def __init__(self, optimizer, steps, patience=5, decreasing=1e-3, verbose=False):
        super().__init__(optimizer, steps, verbose)
        self.decreasing = decreasing
        self.patience, self.patience_count = patience, 0

    def step(self, loss):
        assert self.optimizer.loss is not None, 'scheduler.step() should be called after optimizer.step()'
        self.steps = self.steps + 1
        if self.steps >= self.max_steps:
            self._continual = False
        if (self.optimizer.last - self.optimizer.loss) < self.decreasing:
            self.patience_count = self.patience_count + 1
        else:
            self.patience_count = 0
        if self.patience_count >= self.patience:
            self._continual = False
        if hasattr(self.optimizer, 'reject'):
            if self.optimizer.reject_count >= self.optimizer.reject:
                self._continual = False
