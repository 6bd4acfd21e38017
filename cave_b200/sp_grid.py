"""
Shortest path on an n x n grid without a MIP solver: exact solutions by dynamic programming, the binding
constraints at the optimum in the reference's layout, the PyEPO-style synthetic data generator, and the
normalised regret.  Host-side numpy (dataset construction is offline work); it exists so
that BASELINE.json configs[0] — shortest path 5x5, 40 arcs — can be trained and evaluated end to end in this
image (no Gurobi, no PyEPO).

* arcs / flow rows follow cave_b200.synth.sp_instance (right and down arcs, node-arc incidence F);
* binding constraints at a 0/1 vertex follow src/dataset.py:147-215 of the reference: the 25 flow equalities as
  +-rows (dataset.py:182-184), then -e_k for x_k = 0 and +e_k for x_k = 1 (dataset.py:198-211);
* the generator restates pyepo.data.shortestpath.genData [inferred: PyEPO is not in the reference tree; the
  reference calls the sibling pyepo.data.tsp.genData at code_sample.py:20]: costs are a noisy polynomial of a
  random binary mixing of Gaussian features.
"""
from __future__ import annotations

import numpy as np

from . import synth


def grid_arcs(n: int):
    arcs = []
    for i in range(n):
        for j in range(n):
            v = i * n + j
            if j + 1 < n:
                arcs.append((v, v + 1))
            if i + 1 < n:
                arcs.append((v, v + n))
    return arcs


def gen_data(num_data: int, num_feat: int, grid: int = 5, deg: int = 4, noise: float = 0.5, seed: int = 135):
    """(features [N, p], costs [N, d]) — PyEPO's shortest-path generator restated."""
    rnd = np.random.RandomState(seed)
    d = 2 * grid * (grid - 1)
    B = rnd.binomial(1, 0.5, (d, num_feat))
    x = rnd.normal(0, 1, (num_data, num_feat))
    c = (x @ B.T / np.sqrt(num_feat) + 3) ** deg + 1
    c /= 3.5 ** deg
    c *= rnd.uniform(1 - noise, 1 + noise, (num_data, d))
    return x.astype(np.float32), c.astype(np.float32)


def solve(costs: np.ndarray, grid: int = 5):
    """Exact shortest corner-to-corner paths by DP over the grid DAG (any sign of costs).
    Returns (sol [N, d] uint8, obj [N])."""
    costs = np.atleast_2d(np.asarray(costs, dtype=np.float64))
    arcs = grid_arcs(grid)
    into = [[] for _ in range(grid * grid)]
    for k, (a, b) in enumerate(arcs):
        into[b].append((k, a))
    N = costs.shape[0]
    dist = np.full((N, grid * grid), np.inf)
    dist[:, 0] = 0.0
    pred = np.zeros((N, grid * grid), dtype=np.int64)
    for v in range(1, grid * grid):            # node ids are already a topological order
        cand = np.stack([dist[:, a] + costs[:, k] for k, a in into[v]], axis=1)
        best = np.argmin(cand, axis=1)
        dist[:, v] = cand[np.arange(N), best]
        pred[:, v] = np.asarray([k for k, _ in into[v]])[best]
    sol = np.zeros((N, len(arcs)), dtype=np.uint8)
    tails = np.asarray([a for a, _ in arcs])
    v = np.full(N, grid * grid - 1)
    for _ in range(2 * (grid - 1)):
        k = pred[np.arange(N), v]
        sol[np.arange(N), k] = 1
        v = tails[k]
    return sol, dist[:, -1]


def binding_constraints(sol: np.ndarray, grid: int = 5) -> synth.SparseInstance:
    """Normals of the constraints binding at the 0/1 vertex `sol`: [F; -F; -e_k (x_k = 0); +e_k (x_k = 1)]."""
    arcs = grid_arcs(grid)
    nodes = grid * grid
    out_arcs = [[] for _ in range(nodes)]
    in_arcs = [[] for _ in range(nodes)]
    for k, (a, b) in enumerate(arcs):
        out_arcs[a].append(k)
        in_arcs[b].append(k)
    blocks = []
    for sgn in (1.0, -1.0):
        for v in range(nodes):
            cs = out_arcs[v] + in_arcs[v]
            vs = [sgn] * len(out_arcs[v]) + [-sgn] * len(in_arcs[v])
            blocks.append((cs, np.asarray(vs, dtype=np.float32)))
    return synth._assemble(len(arcs), blocks, np.asarray(sol, dtype=np.uint8))


def normalised_regret(pred_costs: np.ndarray, true_costs: np.ndarray, grid: int = 5) -> float:
    """sum_i (c_i . w(c_hat_i) - z*_i) / sum_i |z*_i|  (PyEPO's metric, README results tables)."""
    w_hat, _ = solve(pred_costs, grid)
    _, z = solve(true_costs, grid)
    achieved = (true_costs.astype(np.float64) * w_hat).sum(axis=1)
    return float((achieved - z).sum() / np.abs(z).sum())
