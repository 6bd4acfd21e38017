import numpy as np
from torch import nn


class optModule(nn.Module):
    def __init__(self, optmodel, processes=1, solve_ratio=1, reduction="mean", dataset=None):
        super().__init__()
        self.optmodel = optmodel
        self.processes = processes
        self.pool = None
        self.solve_ratio = solve_ratio
        self.reduction = reduction
        self._branch_rng = np.random.RandomState()

    def _reduce(self, loss):
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        if self.reduction == "none":
            return loss
        raise ValueError(f"No reduction '{self.reduction}'.")
