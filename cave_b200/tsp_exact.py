"""
TSP without a MIP solver (BASELINE.json configs[1]): the PyEPO-style synthetic data generator,
exact tours by Held-Karp dynamic programming on the GPU (``cave_tsp_solve``, n <= 20 nodes), the binding constraints at
an optimal tour in the reference's layout, and the normalised decision regret.  Evaluation / dataset construction only:
nothing here is on the hot path.

* ``gen_data`` restates ``pyepo.data.tsp.genData`` [inferred: PyEPO is not in the reference tree; the reference calls it
  at code_sample.py:20]: Euclidean distances of random node coordinates plus a noisy polynomial of a random binary
  mixing of Gaussian features.
* ``binding_constraints`` follows src/dataset.py:147-215: the n degree equalities as +-rows (dataset.py:182-184), the
  tracked lazy subtour cuts that are tight at the optimum (dataset.py:186-196; Gurobi's branch-and-cut path is not
  reproducible, so a random subset of the tight cuts — contiguous tour segments — stands in for them, as in
  cave_b200.synth), then -e_k for x_k = 0 and +e_k for x_k = 1 (dataset.py:198-211).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, synth


def gen_data(num_data: int, num_feat: int, num_nodes: int = 20, deg: int = 4, noise: float = 0.5, seed: int = 135):
    """(features [N, p] float32, edge costs [N, n(n-1)/2] float32), edges (i<j) in lexicographic order."""
    rnd = np.random.RandomState(seed)
    n, p, m = num_data, num_feat, num_nodes
    coords = np.concatenate((rnd.uniform(-2, 2, (m // 2, 2)), rnd.normal(0, 1, (m - m // 2, 2))))
    diff = coords[:, None, :] - coords[None, :, :]
    dist = np.sqrt((diff ** 2).sum(-1))
    iu = np.triu_indices(m, k=1)
    base = dist[iu] * 3.0                                   # [d]
    d = len(base)
    x = rnd.normal(0, 1, (n, p))
    Bm = rnd.binomial(1, 0.5, (d, p))
    feat = (x @ Bm.T / np.sqrt(p) + 3.0) ** deg / 3.0 ** (deg - 1)
    c = (base[None, :] + feat) * rnd.uniform(1 - noise, 1 + noise, (n, d))
    return x.astype(np.float32), np.round(c, 4).astype(np.float32)


def solve(costs, num_nodes: int = 20, device=None):
    """Exact tours by Held-Karp on the GPU.  Returns (sol [N, d] uint8 edge incidence, obj [N] float64, tours [N, n])."""
    lib = _lib.load()
    c = torch.as_tensor(costs, dtype=torch.float32)
    if c.dim() == 1:
        c = c[None]
    dev = torch.device(device) if device is not None else (c.device if c.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    c = c.to(dev).contiguous()
    N, d = c.shape
    n = num_nodes
    if d != n * (n - 1) // 2:
        raise ValueError(f"expected {n * (n - 1) // 2} edge costs for {n} nodes, got {d}")
    nb = ctypes.c_size_t()
    _lib.check(lib.cave_tsp_scratch_bytes(N, n, ctypes.byref(nb)))
    scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    tours = torch.empty((N, n), dtype=torch.int32, device=dev)
    obj = torch.empty(N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.cave_tsp_solve(ctypes.c_void_p(c.data_ptr()), N, n, ctypes.c_void_p(tours.data_ptr()),
                                      ctypes.c_void_p(obj.data_ptr()), ctypes.c_void_p(scratch.data_ptr()), nb.value,
                                      ctypes.c_void_p(stream)))
    t = tours.cpu().numpy()
    eidx = synth._edge_index(n)
    sol = np.zeros((N, d), dtype=np.uint8)
    rows = np.repeat(np.arange(N), n)
    sol[rows, eidx[t, np.roll(t, -1, axis=1)].reshape(-1)] = 1
    return sol, obj.cpu().numpy(), t


def binding_constraints(tour: np.ndarray, rng: np.random.Generator, max_cuts: int = 8) -> synth.SparseInstance:
    """Binding-constraint normals at the vertex of `tour` (rows [D; -D; k tight subtour cuts; -e_k; +e_k])."""
    return synth.tsp_instance(len(tour), rng, max_cuts, tour=np.asarray(tour))


def normalised_regret(pred_costs: np.ndarray, true_costs: np.ndarray, num_nodes: int = 20, true_obj: np.ndarray | None = None) -> float:
    """sum_i (c_i . w(pred_i) - c_i . w(c_i)) / sum_i c_i . w(c_i)   (pyepo.metric.regret's normalisation)."""
    sol, _, _ = solve(pred_costs, num_nodes)
    if true_obj is None:
        _, true_obj, _ = solve(true_costs, num_nodes)
    got = (sol.astype(np.float64) * true_costs.astype(np.float64)).sum(axis=1)
    return float((got - true_obj).sum() / np.abs(true_obj).sum())
