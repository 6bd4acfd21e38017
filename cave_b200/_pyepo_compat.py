"""
Stand-in for the three PyEPO symbols the reference's loss modules need, used only when PyEPO
itself is not importable (it is not installed in this image).  With PyEPO present the real
classes are used, so the modules in cave_b200.cave remain drop-ins for src/cave.py.

Inferred from the reference's call sites (PyEPO is not in the reference tree):
  pyepo.EPO.MINIMIZE / MAXIMIZE            src/cave.py:62-67
  pyepo.model.opt.optModel                 test/test_func.py:21-29 (only ``modelSense`` is read)
  pyepo.func.abcmodule.optModule           src/cave.py:53 ``__init__(optmodel, processes,
      solve_ratio, reduction)``; attributes ``optmodel``, ``processes``, ``pool``,
      ``_branch_rng`` (src/cave.py:126, 195, 201); method ``_reduce`` (src/cave.py:73).
"""
from __future__ import annotations

try:  # pragma: no cover - PyEPO is absent in the build image
    from pyepo import EPO
    from pyepo.func.abcmodule import optModule
    from pyepo.model.opt import optModel
    HAVE_PYEPO = True
except Exception:
    from enum import Enum

    import numpy as np
    from torch import nn

    HAVE_PYEPO = False

    class EPO(Enum):
        MINIMIZE = 1
        MAXIMIZE = -1

    class optModel:  # noqa: N801 - reference naming
        """Minimal optimisation-model base: the loss only reads ``modelSense``."""
        modelSense = None

    class optModule(nn.Module):  # noqa: N801 - reference naming
        def __init__(self, optmodel, processes: int = 1, solve_ratio: float = 1.0,
                     reduction: str = "mean", dataset=None) -> None:
            super().__init__()
            if reduction not in ("mean", "sum", "none"):
                raise ValueError(f"No reduction '{reduction}'.")
            self.optmodel = optmodel
            self.processes = processes
            self.pool = None                    # batched backend: no process pool (src/cave.py:242-244)
            self.solve_ratio = solve_ratio
            self.reduction = reduction
            self._branch_rng = np.random.RandomState()

        def _reduce(self, loss):
            if self.reduction == "mean":
                return loss.mean()
            if self.reduction == "sum":
                return loss.sum()
            return loss
