"""ctypes wrapper around the CPU simulation of the solver core (test infrastructure)."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as _build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def forward_backward(pred, ctrs, minimize=True, mode=0, inner_ratio=0.2, reduction="mean", compute_f32=False,
                     force_path=0):
    A = np.ascontiguousarray(ctrs, dtype=np.float32)
    B, m, d = A.shape
    p = np.ascontiguousarray(pred, dtype=np.float64)
    sign = -1.0 if minimize else 1.0
    gscale = 1.0 / B if reduction == "mean" else 1.0
    loss_i = np.zeros(B); grad = np.zeros((B, d)); proj = np.zeros((B, d)); rnorm = np.zeros(B)
    status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    P = ctypes.c_void_p
    lib().hostsim_forward_backward(
        A.ctypes.data_as(P), B, m, d, p.ctypes.data_as(P), ctypes.c_double(sign), mode, ctypes.c_double(inner_ratio),
        ctypes.c_double(gscale), int(compute_f32), int(force_path), loss_i.ctypes.data_as(P), grad.ctypes.data_as(P),
        proj.ctypes.data_as(P), rnorm.ctypes.data_as(P), status.ctypes.data_as(P), iters.ctypes.data_as(P))
    loss = loss_i.mean() if reduction == "mean" else loss_i.sum() if reduction == "sum" else loss_i
    return dict(loss=loss, loss_i=loss_i, grad=grad, proj=proj, rnorm=rnorm, status=status, iters=iters)
