"""Shortest-path utilities (exact DP solver, binding constraints at the vertex, regret): CPU tests."""
import itertools

import numpy as np

from cave_b200 import sp_grid, synth
from oracle import cave_oracle as O


def _all_paths(grid):
    arcs = {a: k for k, a in enumerate(sp_grid.grid_arcs(grid))}
    for moves in set(itertools.permutations([0] * (grid - 1) + [1] * (grid - 1))):
        v, ks = 0, []
        for mv in moves:
            w = v + 1 if mv == 0 else v + grid
            ks.append(arcs[(v, w)])
            v = w
        yield ks


def test_dp_solver_is_exact_against_enumeration():
    rng = np.random.default_rng(0)
    costs = rng.standard_normal((20, 24))            # 4x4 grid, signed costs
    sol, obj = sp_grid.solve(costs, grid=4)
    best = np.min([[c[ks].sum() for ks in _all_paths(4)] for c in costs], axis=1)
    np.testing.assert_allclose(obj, best, atol=1e-12)
    np.testing.assert_allclose((costs * sol).sum(axis=1), obj, atol=1e-12)
    assert (sol.sum(axis=1) == 6).all()


def test_binding_constraints_match_the_synthetic_layout_and_certify_optimality():
    x, c = sp_grid.gen_data(8, 5, seed=1)
    sols, obj = sp_grid.solve(c)
    for s, cost in zip(sols, c):
        inst = sp_grid.binding_constraints(s)
        assert (inst.m, inst.d) == (90, 40)
        A = inst.dense().astype(np.float64)
        # the vertex is optimal for `cost`  <=>  -cost lies in the cone of binding normals (rnorm = 0)
        _, rnorm = O.project_nnls(-cost.astype(np.float64), A, fp64_out=True)
        assert rnorm < 1e-9
    rng = np.random.default_rng(3)
    ref = synth.sp_instance(5, rng)
    mine = sp_grid.binding_constraints(ref.sol)
    np.testing.assert_array_equal(ref.dense(), mine.dense())


def test_regret_is_zero_for_true_costs_and_positive_for_noise():
    x, c = sp_grid.gen_data(64, 5, seed=2)
    assert sp_grid.normalised_regret(c, c) == 0.0
    assert sp_grid.normalised_regret(np.random.default_rng(0).random(c.shape), c) > 0.05
