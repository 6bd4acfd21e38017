"""CPU tests of the round-2 host-side pieces: sparse constraint containers, the TSP data / binding-constraint helpers,
the truncated interior-point emulation of the reference's Clarabel branch (against scipy at convergence), the NUMA
helper, and the layout of the options struct shared with the C ABI."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_sparse_constraints_from_dense_and_from_instances_agree():
    from cave_b200 import SparseConstraints, synth
    insts = synth.make_batch("vrp20", 5, seed=3)
    sc = SparseConstraints.from_instances(insts)
    sd = SparseConstraints.from_dense(synth.densify(insts))
    for a, b in ((sc.inst_off, sd.inst_off), (sc.row_ptr, sd.row_ptr), (sc.col, sd.col), (sc.val, sd.val)):
        assert torch.equal(a, b)
    assert sc.batch == 5 and sc.d == insts[0].d and sc.m_max == max(i.m for i in insts)
    assert int(sc.inst_off[-1]) == sum(i.m for i in insts) and int(sc.row_ptr[-1]) == sum(len(i.vals) for i in insts)
    # columns ascending inside every row (the packer's requirement)
    rp, col = sc.row_ptr.numpy(), sc.col.numpy()
    assert all((np.diff(col[rp[r]:rp[r + 1]]) > 0).all() for r in range(len(rp) - 1))
    assert sc.nbytes() < synth.densify(insts).numel() * 4 // 8


def test_tsp_data_and_binding_constraints_at_a_given_tour():
    from cave_b200 import synth, tsp_exact
    x, c = tsp_exact.gen_data(6, 10, 20, 4, 0.5, seed=1)
    assert x.shape == (6, 10) and c.shape == (6, 190) and (c > 0).all()
    tour = np.random.default_rng(0).permutation(20)
    inst = tsp_exact.binding_constraints(tour, np.random.default_rng(1), 8)
    A = inst.dense()
    eidx = synth._edge_index(20)
    sol = np.zeros(190)
    sol[eidx[tour, np.roll(tour, -1)]] = 1
    assert (inst.sol == sol).all() and inst.m >= 2 * 20 + 190
    assert (A[:20] @ sol == 2).all() and (A[20:40] @ sol == -2).all()            # degree rows are tight at the tour
    ncut = inst.m - 40 - 190
    for r in range(40, 40 + ncut):                                                # subtour cuts: sum_{e in S} x_e = |S| - 1
        k = int(round((1 + np.sqrt(1 + 8 * A[r].sum())) / 2))
        assert A[r] @ sol == k - 1
    assert (np.abs(A[40 + ncut:]).sum(axis=1) == 1).all()                         # bound rows are singletons


def test_truncated_interior_point_emulation_converges_to_the_nnls_projection():
    import scipy.optimize as so
    from clarabel_emulation import project_ipm_truncated
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 14, 9)).astype(np.float32)
    A[1, 10:] = 0                                                                 # padding rows
    c = rng.standard_normal((3, 9)).astype(np.float32)
    full = project_ipm_truncated(torch.tensor(A), torch.tensor(c), 60).numpy()
    three = project_ipm_truncated(torch.tensor(A), torch.tensor(c), 3).numpy()
    for b in range(3):
        rows = np.abs(A[b]).sum(1) > 0
        lam, _ = so.nnls(A[b][rows].astype(np.float64).T, c[b].astype(np.float64))
        ref = A[b][rows].astype(np.float64).T @ lam
        assert np.abs(full[b] - ref).max() < 1e-9
        assert 0 < np.abs(three[b] - ref).max() < 0.5                             # truncated: near, not at, the projection


def test_numa_helper_is_best_effort_and_options_struct_matches_the_header():
    from cave_b200 import _lib
    from cave_b200.parallel import bind_to_gpu_numa_node
    info = bind_to_gpu_numa_node(0)
    assert isinstance(info, dict) and "bound" in info
    # include/cave_b200.h: 2 x int32, double, 2 x int64, 2 x int32, pointer, 2 x int64
    assert ctypes.sizeof(_lib.SolverOpts) == 64
    assert _lib.SolverOpts.dense_mode.offset == 36 and _lib.SolverOpts.inst_index.offset == 40 and _lib.SolverOpts.dense_slots.offset == 56
