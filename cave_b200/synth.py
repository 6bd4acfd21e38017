"""
Synthetic binding-constraint generators (no Gurobi needed).

They reproduce the *shape and values* of the matrices that the reference's
``_extract_tight_normals`` (/root/reference/src/dataset.py:147-215) stores per
instance, following SURVEY.md App. B:

  row order = tight ``<=`` rows, negated ``>=`` rows, ``=`` rows, negated ``=``
  rows (dataset.py:178-184); tracked lazy cuts tight at the optimum
  (dataset.py:186-196); ``-e_k`` rows for ``x_k = 0`` ascending k
  (dataset.py:203-206); ``+e_k`` rows for ``x_k = 1`` (dataset.py:208-211).

Every instance is returned sparse (COO); ``densify`` produces the zero padded
``[B, m_max, d]`` float32 tensor that ``collate_fn`` (dataset.py:133-144) hands
to the loss.  Used by tests/, bench.py and __graft_entry__.smoke() only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SparseInstance:
    """One instance's binding-constraint normals in COO form."""
    m: int              # number of valid rows
    d: int              # number of cost coefficients
    rows: np.ndarray    # int32 [nnz]
    cols: np.ndarray    # int32 [nnz]
    vals: np.ndarray    # float32 [nnz]
    sol: np.ndarray     # uint8 [d]  the optimal vertex the rows are binding at

    def dense(self, m_pad: int | None = None) -> np.ndarray:
        m_pad = self.m if m_pad is None else m_pad
        out = np.zeros((m_pad, self.d), dtype=np.float32)
        out[self.rows, self.cols] = self.vals
        return out


def _edge_index(n: int) -> np.ndarray:
    """idx[i, j] = position of undirected edge (i<j) in lexicographic order."""
    idx = -np.ones((n, n), dtype=np.int64)
    iu = np.triu_indices(n, k=1)
    idx[iu] = np.arange(len(iu[0]))
    idx = np.maximum(idx, idx.T)
    return idx


def _assemble(d, blocks, sol) -> SparseInstance:
    """blocks: list of (list_of_rows) where a row = (cols, vals)."""
    rows, cols, vals = [], [], []
    r = 0
    for cs, vs in blocks:
        cs = np.asarray(cs, dtype=np.int32)
        rows.append(np.full(len(cs), r, dtype=np.int32))
        cols.append(cs)
        vals.append(np.asarray(vs, dtype=np.float32) * np.ones(len(cs), dtype=np.float32))
        r += 1
    # bound rows: -e_k for sol == 0 (ascending k), +e_k for sol == 1
    low = np.where(sol == 0)[0].astype(np.int32)
    high = np.where(sol == 1)[0].astype(np.int32)
    rows.append(np.arange(r, r + len(low), dtype=np.int32)); cols.append(low)
    vals.append(-np.ones(len(low), dtype=np.float32)); r += len(low)
    rows.append(np.arange(r, r + len(high), dtype=np.int32)); cols.append(high)
    vals.append(np.ones(len(high), dtype=np.float32)); r += len(high)
    return SparseInstance(r, d, np.concatenate(rows), np.concatenate(cols),
                          np.concatenate(vals), sol.astype(np.uint8))


def tsp_instance(n: int, rng: np.random.Generator, max_cuts: int, tour: np.ndarray | None = None) -> SparseInstance:
    """TSP-n DFJ: rows [D; -D; k tight subtour cuts; -e_k; +e_k], m = 2n + d + k.  `tour`: the optimal tour the rows are
    binding at (a random permutation when not given)."""
    d = n * (n - 1) // 2
    eidx = _edge_index(n)
    if tour is None:
        tour = rng.permutation(n)
    sol = np.zeros(d, dtype=np.uint8)
    for a in range(n):
        sol[eidx[tour[a], tour[(a + 1) % n]]] = 1
    blocks = []
    inc = [eidx[v, np.arange(n) != v] for v in range(n)]
    for v in range(n):
        blocks.append((inc[v], 1.0))
    for v in range(n):
        blocks.append((inc[v], -1.0))
    k = int(rng.integers(0, max_cuts + 1))
    for _ in range(k):
        # a tight subtour cut: contiguous tour segment S, 2 <= |S| <= n-2
        size = int(rng.integers(2, n - 1))
        start = int(rng.integers(0, n))
        seg = np.sort(tour[(start + np.arange(size)) % n])
        iu = np.triu_indices(size, k=1)
        blocks.append((eidx[seg[iu[0]], seg[iu[1]]], 1.0))
    return _assemble(d, blocks, sol)


def sp_instance(grid: int, rng: np.random.Generator) -> SparseInstance:
    """Shortest path on a grid x grid lattice: rows [F; -F; -e_k; +e_k]."""
    nodes = grid * grid
    arcs = []
    for i in range(grid):
        for j in range(grid):
            v = i * grid + j
            if j + 1 < grid:
                arcs.append((v, v + 1))
            if i + 1 < grid:
                arcs.append((v, v + grid))
    d = len(arcs)
    amap = {a: k for k, a in enumerate(arcs)}
    # random monotone path from corner to corner
    moves = np.array([0] * (grid - 1) + [1] * (grid - 1))
    rng.shuffle(moves)
    sol = np.zeros(d, dtype=np.uint8)
    v = 0
    for mv in moves:
        w = v + 1 if mv == 0 else v + grid
        sol[amap[(v, w)]] = 1
        v = w
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
    return _assemble(d, blocks, sol)


def vrp_instance(n_nodes: int, rng: np.random.Generator, max_cuts: int,
                 n_vehicle: int = 4) -> SparseInstance:
    """CVRP with depot 0: rows [depot (prob 1/2); +/- customer degree; k capacity cuts; bounds]."""
    d = n_nodes * (n_nodes - 1) // 2
    eidx = _edge_index(n_nodes)
    cust = rng.permutation(np.arange(1, n_nodes))
    # split customers into n_vehicle routes with >= 2 customers each
    cut_pts = np.sort(rng.choice(np.arange(2, len(cust) - 1, 2), size=n_vehicle - 1, replace=False))
    routes = np.split(cust, cut_pts)
    sol = np.zeros(d, dtype=np.uint8)
    for r in routes:
        path = np.concatenate([[0], r, [0]])
        for a in range(len(path) - 1):
            sol[eidx[path[a], path[a + 1]]] = 1
    blocks = []
    inc = [eidx[v, np.arange(n_nodes) != v] for v in range(n_nodes)]
    if rng.random() < 0.5:
        blocks.append((inc[0], 1.0))
    for v in range(1, n_nodes):
        blocks.append((inc[v], 1.0))
    for v in range(1, n_nodes):
        blocks.append((inc[v], -1.0))
    k = int(rng.integers(0, max_cuts + 1))
    for _ in range(k):
        # tight capacity cut: a contiguous piece (>= 2 customers) of one route
        r = routes[int(rng.integers(0, len(routes)))]
        size = int(rng.integers(2, len(r) + 1))
        start = int(rng.integers(0, len(r) - size + 1))
        seg = np.sort(r[start:start + size])
        iu = np.triu_indices(size, k=1)
        blocks.append((eidx[seg[iu[0]], seg[iu[1]]], 1.0))
    return _assemble(d, blocks, sol)


def make_batch(kind: str, batch: int, seed: int = 0, max_cuts: int | None = None):
    """kind in {'sp5','tsp20','tsp50','tsp100','vrp20', 'tspN'}; returns list[SparseInstance]."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(batch):
        if kind == "sp5":
            out.append(sp_instance(5, rng))
        elif kind.startswith("tsp"):
            n = int(kind[3:])
            mc = max_cuts if max_cuts is not None else {20: 8, 50: 12}.get(n, 12)
            out.append(tsp_instance(n, rng, mc))
        elif kind == "vrp20":
            out.append(vrp_instance(21, rng, 16 if max_cuts is None else max_cuts))
        else:
            raise ValueError(kind)
    return out


def densify(insts, m_pad: int | None = None, device="cpu", chunk: int = 256):
    """Zero padded float32 [B, m_max, d] torch tensor (the collate_fn layout), built on `device`."""
    import torch
    B, d = len(insts), insts[0].d
    m_max = max(i.m for i in insts) if m_pad is None else m_pad
    out = torch.zeros((B, m_max, d), dtype=torch.float32, device=device)
    flat = out.view(-1)
    for s in range(0, B, chunk):
        idx, val = [], []
        for b in range(s, min(B, s + chunk)):
            it = insts[b]
            idx.append((b * m_max + it.rows.astype(np.int64)) * d + it.cols)
            val.append(it.vals)
        idx = torch.from_numpy(np.concatenate(idx)).to(device)
        val = torch.from_numpy(np.concatenate(val)).to(device)
        flat[idx] = val
    return out


def predictions(insts, seed: int = 0, regime: str = "uniform") -> np.ndarray:
    """c_pred [B, d] float32.  'uniform': U(0,1)^d (SURVEY §6 regime);
    'near': noisy costs for which the stored vertex is near optimal, so cut rows become active."""
    rng = np.random.default_rng(seed + 12345)
    B, d = len(insts), insts[0].d
    if regime == "uniform":
        return rng.random((B, d), dtype=np.float32)
    out = np.empty((B, d), dtype=np.float32)
    for b, it in enumerate(insts):
        c_true = np.where(it.sol > 0, 0.2, 1.0) * (0.5 + rng.random(d))
        out[b] = (c_true * (1.0 + 0.5 * rng.standard_normal(d))).astype(np.float32)
    return out


def dense_batch(batch: int, m: int, d: int, seed: int = 0, device="cpu"):
    """Dense sweep: A ~ N(0,1)^{m x d}, c_pred ~ N(0,1)^d."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    A = torch.randn((batch, m, d), generator=g, device=device, dtype=torch.float32)
    c = torch.randn((batch, d), generator=g, device=device, dtype=torch.float32)
    return A, c
