"""GPU parity tests: the CUDA path, called through the C ABI (cave_b200.qpsolver -> ctypes ->
libcave_b200.so), against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): fp64 compute mode 1e-5 relative, fp32 compute mode 1e-3
relative, on projection, rnorm, loss and gradient."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_names
from oracle import cave_oracle as O

pytestmark = pytest.mark.gpu

RTOL = {"fp64": 1e-5, "fp32": 1e-3}


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _run(pred, ctrs, minimize=True, mode=0, inner_ratio=0.2, reduction="mean", precision="fp64", io64=True, **kw):
    from cave_b200 import cave_forward_backward
    dev = _cuda()
    p = torch.as_tensor(pred, dtype=torch.float64 if io64 else torch.float32, device=dev)
    A = torch.as_tensor(ctrs, dtype=torch.float32, device=dev)
    out = cave_forward_backward(p, A, -1.0 if minimize else 1.0, mode, inner_ratio, reduction, precision=precision,
                                want_proj=mode != 2, want_status=True, **kw)
    torch.cuda.synchronize()
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def _check(out, ref, rtol, mode):
    # gradient entries scale like 1/||c||; a loss of ~0 has a gradient of pure rounding noise
    scale_g = max(np.abs(ref["grad"]).max(), 1e-4)
    np.testing.assert_allclose(out["loss"], ref["loss"], rtol=rtol, atol=rtol * 1e-1)
    np.testing.assert_allclose(out["grad"], ref["grad"], rtol=rtol, atol=rtol * scale_g)
    if mode != 2:
        scale_p = max(np.abs(ref["proj"]).max(), 1e-30)
        np.testing.assert_allclose(out["proj"], ref["proj"], rtol=rtol, atol=rtol * scale_p)
        np.testing.assert_allclose(out["rnorm"], ref["rnorm"], rtol=rtol, atol=rtol * max(ref["rnorm"].max(), 1e-6))
        assert set((out["status"] & 0xff).tolist()) <= {0, 4}, f"solver status {set(out['status'].tolist())}"


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", golden_names())
def test_golden_cases_match_oracle(name, precision):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(minimize=bool(z["minimize"]), mode=int(z["mode"]), inner_ratio=float(z["inner_ratio"]),
              reduction=str(z["reduction"]))
    # precision="fp32" is a float32 FACTOR with float64 iterates and residuals (iterative refinement), so the
    # `rnorm < 1e-7` inside-the-cone test of the push-inside step (src/cave.py:218) resolves in both modes: no instance
    # is left out, the inside_inner golden (every instance inside its cone) included.
    ref = O.forward_backward(z["pred"], z["ctrs"], fp64=True, **kw)
    out = _run(z["pred"], z["ctrs"], precision=precision, **kw)
    if kw["mode"] == 1:
        inside_ref = ref["rnorm"] < 1e-7
        assert ((out["rnorm"] < 1e-7) == inside_ref).all(), (out["rnorm"], ref["rnorm"])
    _check(out, ref, RTOL[precision], kw["mode"])


@pytest.mark.parametrize("name", ["tf_randn_exact", "tsp20_inner_near", "sp5_exact_uniform"])
def test_matches_reference_golden_float32_io(name):
    """float32 in/out like the reference: compare with the reference's own stored outputs."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(minimize=bool(z["minimize"]), mode=int(z["mode"]), inner_ratio=float(z["inner_ratio"]),
              reduction=str(z["reduction"]))
    out = _run(z["pred"], z["ctrs"], io64=False, **kw)
    np.testing.assert_allclose(out["loss"], z["loss"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out["proj"], z["proj"], rtol=1e-5, atol=1e-5 * np.abs(z["proj"]).max())
    np.testing.assert_allclose(out["grad"], z["grad"], rtol=1e-3, atol=1e-5 * max(np.abs(z["grad"]).max(), 1e-30))


@pytest.mark.parametrize("kind,batch,mode", [("sp5", 32, 0), ("tsp20", 32, 1), ("vrp20", 32, 1), ("tsp50", 3, 1)])
@pytest.mark.parametrize("regime", ["uniform", "near"])
def test_structured_configs(kind, batch, mode, regime):
    from cave_b200 import synth
    insts = synth.make_batch(kind, batch, seed=11)
    ctrs = synth.densify(insts).numpy()
    pred = synth.predictions(insts, 11, regime)
    ref = O.forward_backward(pred, ctrs, mode=mode, fp64=True)
    out = _run(pred, ctrs, mode=mode)
    _check(out, ref, 1e-5, mode)


@pytest.mark.parametrize("m,d,batch", [(15, 10, 32), (5, 8, 16), (64, 190, 8), (256, 190, 4), (300, 100, 4), (1024, 190, 2)])
def test_dense_sweep_shapes(m, d, batch):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((batch, m, d)).astype(np.float32)
    c = rng.standard_normal((batch, d)).astype(np.float32)
    ref = O.forward_backward(c, A, mode=0, fp64=True)
    out = _run(c, A, mode=0)
    _check(out, ref, 1e-5, 0)
    assert (out["status"] & 0x100).all()        # pure general rows -> Lawson-Hanson path


def test_kat_identity_and_empty_and_zero():
    rng = np.random.default_rng(1)
    c = rng.standard_normal((3, 7))
    A = np.broadcast_to(np.eye(7, dtype=np.float32), (3, 7, 7)).copy()
    out = _run(-c, A, mode=0)                       # minimize: c = -pred
    np.testing.assert_allclose(out["proj"], np.maximum(c, 0.0), atol=1e-12)
    out = _run(-c, np.zeros((3, 4, 7), np.float32), mode=1, reduction="none")
    np.testing.assert_allclose(out["proj"], c, atol=0)     # empty cone returns c (src/cave.py:304-305)
    np.testing.assert_allclose(out["loss"], 0.0, atol=1e-12)
    out = _run(np.zeros((2, 6)), rng.random((2, 3, 6)).astype(np.float32), mode=0)
    assert out["loss"] == pytest.approx(1.0, abs=1e-12) and np.abs(out["grad"]).max() == 0.0


def test_moreau_kkt_on_device_result():
    from cave_b200 import synth
    insts = synth.make_batch("tsp20", 8, seed=3)
    ctrs = synth.densify(insts).numpy()
    pred = synth.predictions(insts, 3, "near")
    out = _run(pred, ctrs, mode=0)
    for b in range(8):
        c = -pred[b].astype(np.float64)
        q = c - out["proj"][b]
        A = ctrs[b].astype(np.float64)
        assert (A @ q).max() <= 1e-8 * np.abs(c).max() * 50
        assert abs(out["proj"][b] @ q) <= 1e-8 * (c @ c)


def test_padding_and_row_count_hint_change_nothing():
    from cave_b200 import synth
    insts = synth.make_batch("vrp20", 8, seed=5)
    pred = synth.predictions(insts, 5, "near")
    a = _run(pred, synth.densify(insts).numpy(), mode=1, reduction="none")
    m_pad = max(i.m for i in insts) + 9
    b = _run(pred, synth.densify(insts, m_pad=m_pad).numpy(), mode=1, reduction="none")
    np.testing.assert_array_equal(a["loss"], b["loss"])
    np.testing.assert_array_equal(a["grad"], b["grad"])
    m_rows = torch.tensor([i.m for i in insts], dtype=torch.int32)
    c = _run(pred, synth.densify(insts, m_pad=m_pad).numpy(), mode=1, reduction="none", m_rows=m_rows)
    np.testing.assert_array_equal(a["loss"], c["loss"])


def test_bitwise_reproducible():
    from cave_b200 import synth
    insts = synth.make_batch("tsp20", 16, seed=9)
    ctrs, pred = synth.densify(insts).numpy(), synth.predictions(insts, 9, "uniform")
    a, b = _run(pred, ctrs, mode=1), _run(pred, ctrs, mode=1)
    np.testing.assert_array_equal(a["grad"], b["grad"])
    np.testing.assert_array_equal(a["loss"], b["loss"])


def test_modules_forward_backward_match_oracle_and_autograd():
    from cave_b200 import EPO, exactConeAlignedCosine, innerConeAlignedCosine, synth
    dev = _cuda()

    class M:
        modelSense = EPO.MINIMIZE

    insts = synth.make_batch("tsp20", 8, seed=2)
    ctrs = synth.densify(insts, device=dev)
    pred_np = synth.predictions(insts, 2, "near")
    for cls, mode, kw in ((exactConeAlignedCosine, 0, {}), (innerConeAlignedCosine, 1, dict(seed=0, inner_ratio=0.3))):
        for reduction in ("mean", "sum", "none"):
            pred = torch.tensor(pred_np, device=dev, requires_grad=True)
            loss = cls(M(), solver="cuda", reduction=reduction, **kw)(pred, ctrs)
            w = torch.linspace(0.5, 1.5, loss.numel(), device=dev).reshape(loss.shape)
            (loss * w).sum().backward()
            ref = O.forward_backward(pred_np, ctrs.cpu().numpy(), mode=mode, inner_ratio=kw.get("inner_ratio", 0.2),
                                     reduction=reduction, fp64=True)
            np.testing.assert_allclose(loss.detach().cpu().numpy(), ref["loss"], rtol=1e-4, atol=1e-6)
            wg = w.cpu().numpy().reshape(-1, 1) if reduction == "none" else float(w)
            np.testing.assert_allclose(pred.grad.cpu().numpy(), ref["grad"] * wg, rtol=1e-3,
                                       atol=1e-5 * np.abs(ref["grad"]).max())


def test_host_tensors_roundtrip_and_project_cuda():
    from cave_b200 import project_cuda
    _cuda()
    torch.manual_seed(1)
    costs, bctrs = torch.randn(8, 10), torch.randn(8, 15, 10)
    proj, rnorm = project_cuda(bctrs, -costs)
    assert proj.device.type == "cpu" and proj.dtype == torch.float32
    ref_p, ref_r = O.batch_project((-costs).numpy(), bctrs.numpy(), fp64=True)
    np.testing.assert_allclose(proj.numpy(), ref_p, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rnorm.numpy(), ref_r, rtol=1e-5, atol=1e-6)


def test_full_size_tsp50_properties():
    """BASELINE config 3 shape at a reduced batch that still fills the GPU: size-independent checks
    (Moreau decomposition, cone membership of the residual, idempotence of the projection)."""
    from cave_b200 import cave_forward_backward, project_cuda, synth
    dev = _cuda()
    insts = synth.make_batch("tsp50", 296, seed=4)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 4, "near"), dtype=torch.float64, device=dev)
    out = cave_forward_backward(pred, A, -1.0, 1, want_proj=True, want_status=True)
    assert int((out["status"] & 0xff).max()) == 0
    c, p = -pred, out["proj"]
    q = c - p
    Aq = torch.bmm(A.double(), q.unsqueeze(2)).squeeze(2)
    assert float(Aq.max()) <= 1e-8 * float(c.abs().max()) * 50                 # q in the polar cone
    assert float(((p * q).sum(1).abs() / (c * c).sum(1)).max()) <= 1e-9        # <p, q> = 0
    p2, r2 = project_cuda(A, p)
    assert float((p2 - p).abs().max()) <= 1e-7 * float(p.abs().max())          # idempotent (f* = 0: degenerate, converges slowly)
    assert float(r2.max()) <= 1e-7 * float(c.norm(dim=1).max())
    assert torch.isfinite(out["grad"]).all()


def test_scan_kernel_variants_agree(monkeypatch):
    """The warp-streaming scan kernel (small rows) and the tile kernel (any row length) must produce the
    same pack: identical row classes / projections, averages equal up to fp32 summation order."""
    from cave_b200 import synth
    insts = synth.make_batch("vrp20", 12, seed=13)
    ctrs, pred = synth.densify(insts).numpy(), synth.predictions(insts, 13, "near")
    a = _run(pred, ctrs, mode=1, reduction="none")
    monkeypatch.setenv("CAVE_SCAN_KERNEL", "tile")
    b = _run(pred, ctrs, mode=1, reduction="none")
    np.testing.assert_array_equal(a["proj"], b["proj"])
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(a["grad"], b["grad"], rtol=1e-5, atol=1e-9)


def test_long_rows_use_tile_kernel_and_fallback_csr():
    """d = 3000: rows too long for the warp-streaming scan; dense general rows overflow the packed CSR, so
    the solver rebuilds its sparse rows from A (fallback path) — with and without singleton rows."""
    rng = np.random.default_rng(8)
    B, m, d = 3, 40, 3000
    A = rng.standard_normal((B, m, d)).astype(np.float32)
    c = rng.standard_normal((B, d)).astype(np.float32)
    ref = O.forward_backward(c, A, mode=0, fp64=True)
    _check(_run(c, A, mode=0), ref, 1e-5, 0)                       # Lawson-Hanson on dense rows
    A2 = np.concatenate([A[:, :30], np.zeros((B, 600, d), np.float32)], axis=1)
    idx = rng.permutation(d)[:600]
    A2[:, 30 + np.arange(600), idx] = rng.choice([-1.0, 1.0], size=600).astype(np.float32)
    ref = O.forward_backward(c, A2, mode=1, fp64=True)
    out = _run(c, A2, mode=1)
    _check(out, ref, 1e-5, 1)
    assert not (out["status"] & 0x100).any()                       # Newton path, CSR rebuilt from A


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_reproducible_at_scale_and_against_kernel_source_on_cpu(precision):
    """Race detector: 592 TSP-50 instances (two full waves of resident CTAs) three times, bitwise equal;
    and the first instances against the same kernel source compiled for the CPU (tests/hostsim)."""
    import sys
    from cave_b200 import cave_forward_backward, synth
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import sim
    dev = _cuda()
    insts = synth.make_batch("tsp50", 592, seed=21)
    A = synth.densify(insts, device=dev)
    pred_np = synth.predictions(insts, 21, "near")
    pred = torch.tensor(pred_np, dtype=torch.float64, device=dev)
    runs = []
    for _ in range(3):
        o = cave_forward_backward(pred, A, -1.0, 1, 0.2, "none", precision=precision, want_proj=True, want_status=True)
        runs.append({k: v.clone() for k, v in o.items()})
    for k in ("loss", "grad", "proj", "rnorm", "iters"):
        assert torch.equal(runs[0][k], runs[1][k]) and torch.equal(runs[0][k], runs[2][k]), k
    assert int((runs[0]["status"] & 0xff).max()) == 0
    n = 6
    ref = sim.forward_backward(pred_np[:n], A[:n].cpu().numpy(), mode=1, inner_ratio=0.2, reduction="none",
                               compute_f32=precision == "fp32")
    np.testing.assert_allclose(runs[0]["proj"][:n].cpu().numpy(), ref["proj"], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(runs[0]["iters"][:n].cpu().numpy(), ref["iters"])


def test_device_resident_dataset_index_matches_dense_batches():
    """§8f-1: pack the whole dataset once, then address batches by instance index (no dense batch tensor, no
    scan pass).  Must equal the dense-batch call on the gathered rows, with and without the dense tensor."""
    from cave_b200 import EPO, cave_forward_backward, innerConeAlignedCosine, pack_constraints, synth
    dev = _cuda()
    insts = synth.make_batch("vrp20", 40, seed=31)
    allc = synth.densify(insts, device=dev)
    pk = pack_constraints(allc)
    idx = torch.tensor([7, 3, 39, 0, 3, 21, 12], dtype=torch.int32, device=dev)
    pred = torch.tensor(synth.predictions(insts, 31, "near"), device=dev)[idx.long()]
    ref = cave_forward_backward(pred, allc[idx.long()].contiguous(), -1.0, 1, 0.2, "none", want_proj=True, want_status=True)
    for ctrs in (allc, None):
        out = cave_forward_backward(pred, ctrs, -1.0, 1, 0.2, "none", want_proj=True, want_status=True, pack=pk, index=idx)
        for k in ("loss", "grad", "proj", "rnorm", "status"):
            assert torch.equal(out[k], ref[k]), k
    pk2 = pack_constraints(allc, keep_dense=False)

    class M:
        modelSense = EPO.MINIMIZE
    p = pred.clone().requires_grad_(True)
    loss = innerConeAlignedCosine(M(), solver="cuda", seed=0, reduction="none")(p, pk2, index=idx)
    loss.sum().backward()
    assert torch.allclose(loss, ref["loss"]) and torch.allclose(p.grad, ref["grad"])
    # dense general rows do not fit the packed CSR: with the dense tensor dropped they must report, not guess
    A = torch.randn(3, 30, 3000, device=dev)
    pkd = pack_constraints(A, keep_dense=False)
    out = cave_forward_backward(torch.randn(2, 3000, device=dev), None, -1.0, 0, want_status=True, pack=pkd,
                                index=torch.tensor([2, 0], dtype=torch.int32, device=dev))
    assert ((out["status"] & 0xff) == 3).all() and torch.isnan(out["loss"]).all()


def test_status_words_report_caps_and_bad_input():
    from cave_b200 import cave_forward_backward, synth
    dev = _cuda()
    insts = synth.make_batch("tsp20", 4, seed=41)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 41, "near"), dtype=torch.float64, device=dev)
    out = cave_forward_backward(pred, A, -1.0, 1, 0.2, "none", want_status=True, max_iter=1)
    assert ((out["status"] & 0xff) == 1).all() and (out["iters"] == 1).all()      # iteration cap, best iterate
    assert torch.isfinite(out["loss"]).all() and torch.isfinite(out["grad"]).all()
    bad = pred.clone()
    bad[2, 5] = float("nan")
    out = cave_forward_backward(bad, A, -1.0, 1, 0.2, "none", want_status=True)
    st = (out["status"] & 0xff).cpu().tolist()
    assert st[2] == 5 and st[0] == 0 and st[1] == 0 and st[3] == 0
    assert torch.isnan(out["loss"][2]) and torch.isfinite(out["loss"][[0, 1, 3]]).all()


def test_training_loop_reduces_the_loss():
    """code_sample.py analogue (linear predictor, CaVE+, Adam) on synthetic TSP-20 with the device-resident
    dataset: the loss must go down — exercises autograd, the index path and repeated calls end to end."""
    import importlib.util
    _cuda()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_tsp20_cave_plus.py")
    spec = importlib.util.spec_from_file_location("train_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist = mod.train(num_data=128, epochs=6, batch=32, verbose=False)
    assert np.isfinite(hist).all() and hist[-1] < 0.6 * hist[0], hist


def test_sp5_end_to_end_regret_config0():
    """BASELINE.json configs[0]: shortest path 5x5, linear predictor, batch 32, CaVE Exact with solver='cuda';
    exact DP solver for the regret.  Training must bring the normalised test regret well below the untrained one."""
    import importlib.util
    _cuda()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_sp5_regret.py")
    spec = importlib.util.spec_from_file_location("sp5_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r0, r1 = mod.run("cave-e", n_train=400, n_test=400, epochs=6, verbose=False)
    assert r1 < 0.5 * r0 and r1 < 0.25, (r0, r1)


@pytest.mark.parametrize("kind,batch,mode", [("tsp20", 24, 1), ("sp5", 40, 0), ("tsp50", 4, 1)])
def test_every_solve_launch_configuration_matches_the_oracle(kind, batch, mode, monkeypatch):
    """The solve kernel is enqueued as 64x8, 128x4, 160x3 and 256x2 (threads x CTAs/SM) and the pack's plan statistics
    pick one on the device.  Forcing each of them (CAVE_SOLVE_CFG) must give the oracle's answer — including the
    small configurations on TSP-50, whose working set then spills to the global slot."""
    from cave_b200 import synth
    insts = synth.make_batch(kind, batch, seed=21)
    ctrs, pred = synth.densify(insts).numpy(), synth.predictions(insts, 21, "near")
    ref = O.forward_backward(pred, ctrs, mode=mode, fp64=True)
    _check(_run(pred, ctrs, mode=mode), ref, 1e-5, mode)              # the configuration the plan selects
    for cfg in range(4):
        monkeypatch.setenv("CAVE_SOLVE_CFG", str(cfg))
        _check(_run(pred, ctrs, mode=mode), ref, 1e-5, mode)


def test_launch_plan_tracks_the_instance_size():
    """Small instances select many small CTAs per SM, TSP-50 a wider CTA; dense (Lawson-Hanson) instances whose
    Gram does not fit the small configurations select the widest one."""
    from cave_b200 import pack_constraints, synth
    dev = _cuda()
    small = pack_constraints(synth.densify(synth.make_batch("tsp20", 16, seed=3)).to(dev)).launch_plan("fp64")
    big = pack_constraints(synth.densify(synth.make_batch("tsp50", 4, seed=3)).to(dev)).launch_plan("fp64")
    assert small["threads"] * small["ctas_per_sm"] >= 480 and big["threads"] * big["ctas_per_sm"] >= 480
    assert small["smem_bytes"] < big["smem_bytes"]
    assert small["avg_work_bytes_f64"] < big["avg_work_bytes_f64"] and small["avg_work_bytes_f32"] <= small["avg_work_bytes_f64"]
    A, _ = synth.dense_batch(4, 256, 190, seed=5, device=dev)
    dense = pack_constraints(A.contiguous()).launch_plan("fp64")
    assert dense["config"] == 3


def test_tsp50_benchmark_regime_sample_matches_oracle():
    """The benchmarked workload itself (TSP-50, uniform predictions, generator seed 1000 as in bench.py): the first instances
    of the timed batch against the oracle, fp64 and fp32-factor modes (VERDICT round 1, weak point 2)."""
    from cave_b200 import synth
    insts = synth.make_batch("tsp50", 8, seed=1000)
    ctrs = synth.densify(insts).numpy()
    pred = synth.predictions(insts, 1000, "uniform")
    ref = O.forward_backward(pred, ctrs, mode=1, inner_ratio=0.2, reduction="none", fp64=True)
    for precision in ("fp64", "fp32"):
        out = _run(pred, ctrs, mode=1, inner_ratio=0.2, reduction="none", precision=precision)
        _check(out, ref, RTOL[precision], 1)


def test_newton_ends_at_the_floating_point_floor_of_a_degenerate_instance():
    """SP 5x5, generator seed 1001, instance 3905, uniform predictions (tests/golden/regress/sp5_newton_floor.npz; found by
    the two-GPU bench, whose rank 1 draws seed 1001): the float64 Newton iteration reaches the solution in 9 steps, then its
    KKT residual hovers at 25 x tol while every further step is accepted with a length of 2^-24 and no change of the objective.
    It used to run to the iteration cap (status ITER_CAP, 5 ms for one 90 x 40 instance); it must stop, converged, and match
    the oracle."""
    from cave_b200 import cave_forward_backward
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regress", "sp5_newton_floor.npz"))
    dev = torch.device("cuda:0")
    A = torch.tensor(z["A"][None], device=dev)
    pred = torch.tensor(z["pred"][None], device=dev)
    for prec in ("fp64", "fp32"):
        out = cave_forward_backward(pred, A, -1.0, 0, 0.0, "none", precision=prec, want_proj=True, want_status=True)
        assert int(out["status"][0]) & 0xff == 0 and int(out["iters"][0]) <= 20, (prec, int(out["status"][0]), int(out["iters"][0]))
        ref_p, ref_r = O.batch_project(-z["pred"][None].astype(np.float64), z["A"][None], fp64=True)
        assert np.abs(out["proj"].double().cpu().numpy() - ref_p).max() <= 1e-6 * np.abs(ref_p).max()
        np.testing.assert_allclose(out["rnorm"].double().cpu().numpy(), ref_r, rtol=1e-6)
