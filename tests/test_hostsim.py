"""CPU logic tests of the KERNEL SOURCE: cave_b200/csrc/solver_core.cuh compiled with
CAVE_HOST_SIM (one thread, warp width 1; tests/hostsim) against the oracle.  This checks the
algorithm the CUDA kernel runs (CSR/CSC build, +- row merging, projected Newton, Lawson-Hanson,
fused epilogue) without a GPU; races and memory-model issues are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, golden_names
from oracle import cave_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import sim  # noqa: E402


def _cmp(out, ref, mode, rtol=1e-6):
    np.testing.assert_allclose(out["loss"], ref["loss"], rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(out["grad"], ref["grad"], rtol=rtol, atol=1e-6 * max(np.abs(ref["grad"]).max(), 1e-4))
    if mode != 2:
        np.testing.assert_allclose(out["proj"], ref["proj"], rtol=rtol, atol=1e-8 * max(np.abs(ref["proj"]).max(), 1e-30))
        np.testing.assert_allclose(out["rnorm"], ref["rnorm"], rtol=rtol, atol=1e-9 * max(ref["rnorm"].max(), 1.0))
        assert set((out["status"] & 0xff).tolist()) <= {0, 4}


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("name", golden_names())
def test_kernel_source_on_golden_cases(name, f32):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(minimize=bool(z["minimize"]), mode=int(z["mode"]), inner_ratio=float(z["inner_ratio"]),
              reduction=str(z["reduction"]))
    ref = O.forward_backward(z["pred"], z["ctrs"], fp64=True, **kw)
    out = sim.forward_backward(z["pred"], z["ctrs"], compute_f32=f32, **kw)
    _cmp(out, ref, kw["mode"])


@pytest.mark.parametrize("kind,regime", [("sp5", "uniform"), ("sp5", "near"), ("tsp20", "uniform"), ("tsp20", "near"),
                                         ("vrp20", "uniform"), ("vrp20", "near")])
def test_kernel_source_structured(kind, regime):
    from cave_b200 import synth
    insts = synth.make_batch(kind, 24, seed=21)
    ctrs, pred = synth.densify(insts).numpy(), synth.predictions(insts, 21, regime)
    for mode in (0, 1):
        ref = O.forward_backward(pred, ctrs, mode=mode, fp64=True)
        _cmp(sim.forward_backward(pred, ctrs, mode=mode), ref, mode)


@pytest.mark.parametrize("m,d,batch", [(15, 10, 32), (5, 8, 16), (40, 12, 16), (64, 190, 4), (300, 100, 2)])
def test_kernel_source_dense_lawson_hanson(m, d, batch):
    rng = np.random.default_rng(17)
    A = rng.standard_normal((batch, m, d)).astype(np.float32)
    c = rng.standard_normal((batch, d)).astype(np.float32)
    ref = O.forward_backward(c, A, mode=0, fp64=True)
    for f32 in (False, True):
        out = sim.forward_backward(c, A, mode=0, compute_f32=f32)
        _cmp(out, ref, 0)
        assert (out["status"] & 0x100).all()


def test_kernel_source_newton_on_dense_rows():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((8, 12, 20)).astype(np.float32)
    c = rng.standard_normal((8, 20)).astype(np.float32)
    ref = O.forward_backward(c, A, mode=0, fp64=True)
    out = sim.forward_backward(c, A, mode=0, force_path=1)
    _cmp(out, ref, 0)
    assert not (out["status"] & 0x100).any()


def test_kernel_source_inside_cone_and_mixed_singletons():
    from cave_b200 import synth
    rng = np.random.default_rng(5)
    insts = synth.make_batch("tsp20", 6, seed=5)
    ctrs = synth.densify(insts).numpy()
    lam = rng.random((6, ctrs.shape[1])).astype(np.float64) + 0.05
    c = np.einsum("bm,bmd->bd", lam, ctrs.astype(np.float64))
    out = sim.forward_backward(-c, ctrs, mode=1, reduction="none")
    assert out["rnorm"].max() < 1e-7 * 10 and np.abs(out["loss"]).max() < 1e-12     # un-pushed target
    # coordinates with both +e_k and -e_k rows, rows with duplicates, a coordinate without singleton
    A = ctrs[:2].copy()
    A[:, -1, :] = 0.0
    A[:, -1, 3] = 2.5
    A[:, -2, :] = 0.0
    A[:, -2, 3] = -0.5
    pred = rng.standard_normal((2, ctrs.shape[2])).astype(np.float32)
    ref = O.forward_backward(pred, A, mode=1, fp64=True)
    _cmp(sim.forward_backward(pred, A, mode=1), ref, 1)
