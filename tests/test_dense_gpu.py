"""Dense regime (BASELINE.json configs[4]): the tensor-core Gram kernel against a float64
product of the same rows, and the Gram-space solver + float64 polish against the oracle (scipy.optimize.nnls, the
reference's call at src/cave.py:307) at the shapes the sweep names.  All through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import cave_oracle as O

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _dense(B, m, d, seed, kind="randn"):
    g = torch.Generator(device=_dev()).manual_seed(seed)
    if kind == "randn":
        A = torch.randn((B, m, d), generator=g, device=_dev())
    else:
        A = torch.rand((B, m, d), generator=g, device=_dev())
    c = torch.randn((B, d), generator=g, device=_dev(), dtype=torch.float64)
    return A, c


@pytest.mark.parametrize("m,d", [(128, 64), (256, 190), (300, 333), (1024, 1225)])
def test_gram_kernel_matches_float64_product(m, d):
    from cave_b200.qpsolver import dense_gram
    A, _ = _dense(3, m, d, seed=m * 31 + d)
    if m > 140:
        A[1, m - 9:] = 0                    # ragged: instance 1 has 9 padding rows
    G, cnt = dense_gram(A)
    assert cnt == 3
    for b in range(3):
        mv = int((A[b].abs().sum(1) > 0).sum())
        ref = A[b, :mv].double() @ A[b, :mv].double().T
        got = G[b, :mv, :mv].double()
        nrm = A[b, :mv].double().norm(dim=1)
        err = ((got - ref).abs() / (nrm[:, None] * nrm[None, :])).max().item()
        # 3xTF32 operands (2^-22) with float32 accumulation in the tensor core (truncating: grows with d)
        assert err < 4e-5, (b, err)
        assert torch.equal(G[b, :mv, :mv], G[b, :mv, :mv].T)          # exactly symmetric by construction
        blk = (mv + 127) // 128 * 128
        assert float(G[b, mv:blk, :blk].abs().max()) == 0.0 if blk > mv else True


@pytest.mark.parametrize("d,m,B", [(1225, 256, 6), (1225, 1024, 3), (4950, 256, 3), (190, 128, 8), (640, 512, 4)])
def test_dense_path_matches_oracle(d, m, B):
    from cave_b200 import _lib, cave_forward_backward
    A, c = _dense(B, m, d, seed=d * 7 + m)
    out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True)     # dense: auto
    st = out["status"].cpu().numpy()
    assert ((st & _lib.ST_PATH_GRAM) != 0).all(), st             # took the tensor-core path
    assert ((st & 0xff) == 0).all(), st
    n = min(B, 3)
    ref_p, ref_r = O.batch_project(c[:n].cpu().numpy(), A[:n].cpu().numpy(), fp64=True)
    got_p = out["proj"][:n].cpu().numpy()
    assert np.abs(got_p - ref_p).max() <= 1e-5 * np.abs(ref_p).max()          # north_star fp64 tolerance (measured ~1e-10)
    np.testing.assert_allclose(out["rnorm"][:n].cpu().numpy(), ref_r, rtol=1e-7, atol=1e-9)
    # and against the Lawson-Hanson path of the general kernel on ALL instances (dense path switched off)
    lh = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=False)
    assert ((lh["status"].cpu().numpy() & _lib.ST_PATH_LH) != 0).all()
    scale = float(lh["proj"].abs().max())
    assert float((out["proj"] - lh["proj"]).abs().max()) <= 1e-8 * scale
    assert float((out["grad"] - lh["grad"]).abs().max()) <= 1e-8 * max(float(lh["grad"].abs().max()), 1e-30)
    assert float((out["loss_i"] - lh["loss_i"]).abs().max()) <= 1e-9


def test_dense_path_optimality_conditions_and_modes():
    """Moreau / KKT on the device result in float64 (SURVEY 8c KAT 7), float32 I/O, the inner push-inside epilogue, and
    positive (rand) matrices whose solutions have tiny supports."""
    from cave_b200 import _lib, cave_forward_backward
    for kind in ("randn", "rand"):
        A, c = _dense(5, 384, 700, seed=11, kind=kind)
        out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True)
        st = out["status"].cpu().numpy()
        assert ((st & _lib.ST_PATH_GRAM) != 0).all() and ((st & 0xff) == 0).all(), st
        p = out["proj"]
        q = c - p
        w = torch.einsum("bmd,bd->bm", A.double(), q)
        cn = c.norm(dim=1)
        an = A.double().norm(dim=2).max(dim=1).values
        assert float((w.max(dim=1).values / (cn * an)).max()) <= 1e-9           # q in the polar cone
        assert float(((p * q).sum(1).abs() / cn ** 2).max()) <= 1e-9           # <p, c - p> = 0
    A, c = _dense(4, 256, 300, seed=5)
    ref = O.forward_backward(c.float().cpu().numpy(), A.cpu().numpy(), minimize=True, mode=O.MODE_INNER, inner_ratio=0.2, fp64=True)
    out = cave_forward_backward(c.float(), A, -1.0, 1, 0.2, "mean", want_status=True)
    assert ((out["status"].cpu().numpy() & _lib.ST_PATH_GRAM) != 0).all()
    np.testing.assert_allclose(float(out["loss"]), float(ref["loss"]), rtol=1e-5)
    np.testing.assert_allclose(out["grad"].cpu().numpy(), ref["grad"], rtol=1e-4, atol=1e-6 * np.abs(ref["grad"]).max())


def test_dense_batch_in_several_rounds_and_mixed_with_structured_instances():
    """More dense instances than workspace slots (rounds), instances that are not dense in the same batch (taken by
    the general kernel), and an instance with a NaN prediction."""
    from cave_b200 import _lib, cave_forward_backward
    A, c = _dense(7, 160, 200, seed=3)
    A[2, :, :] = 0
    A[2, :40, :40] = torch.eye(40, device=_dev())       # singleton rows only: closed form in the general kernel
    A[5, 100:, :] = 0                                    # 100 valid rows: below the dense threshold -> Lawson-Hanson
    c[6, 3] = float("nan")
    one = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=False)
    out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense_slots=2)
    st = out["status"].cpu().numpy()
    assert [(int(s) & _lib.ST_PATH_GRAM) != 0 for s in st] == [True, True, False, True, True, False, True]
    assert (st[6] & 0xff) == _lib.ST_BADINPUT and bool(torch.isnan(out["grad"][6]).all())
    ok = [0, 1, 2, 3, 4, 5]
    scale = float(one["proj"][ok].abs().max())
    assert float((out["proj"][ok] - one["proj"][ok]).abs().max()) <= 1e-8 * scale
    assert float((out["loss_i"][ok] - one["loss_i"][ok]).abs().max()) <= 1e-9


@pytest.mark.parametrize("d,m,B", [(1225, 1024, 3), (640, 512, 4), (333, 300, 5)])
def test_tensor_core_cholesky_update_matches_ffma_update_and_oracle(d, m, B, monkeypatch):
    """CAVE_DENSE_TC=1 runs the block-column update of the dense Cholesky on the tensor cores (tcgen05, 3 x TF32, accumulators
    in TMEM) instead of the FFMA micro-kernel.  The factor only preconditions the Newton steps (the polish phase works on A in
    float64), so both variants must reach the same projection, and the oracle's."""
    from cave_b200 import _lib, cave_forward_backward
    A, c = _dense(B, m, d, seed=d * 3 + m)
    A[B - 1, m - 37:] = 0                                   # ragged: 37 padding rows (block column with < 64 rows)
    monkeypatch.setenv("CAVE_DENSE_TC", "0")
    ref = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=True)
    monkeypatch.setenv("CAVE_DENSE_TC", "1")
    out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=True)
    st = out["status"].cpu().numpy()
    assert ((st & _lib.ST_PATH_GRAM) != 0).all() and ((st & 0xff) == 0).all(), st
    scale = float(ref["proj"].abs().max())
    assert float((out["proj"] - ref["proj"]).abs().max()) <= 1e-8 * scale
    assert float((out["grad"] - ref["grad"]).abs().max()) <= 1e-8 * max(float(ref["grad"].abs().max()), 1e-30)
    ref_p, ref_r = O.batch_project(c[:2].cpu().numpy(), A[:2].cpu().numpy(), fp64=True)
    assert np.abs(out["proj"][:2].cpu().numpy() - ref_p).max() <= 1e-5 * np.abs(ref_p).max()
    np.testing.assert_allclose(out["rnorm"][:2].cpu().numpy(), ref_r, rtol=1e-7, atol=1e-9)
