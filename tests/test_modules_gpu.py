"""GPU tests of the module-level contract that the reference's own tests exercise (test/test_func.py:285-292 calls
``_get_projection`` directly; :216-227 checks seeded branch sequences): the two-step ``_get_projection`` methods
against the oracle's targets, and a CaVE Hybrid (solve_ratio = 0.3, BASELINE.json configs[3]) multi-call sequence
against the oracle driven by the same ``RandomState``."""
import numpy as np
import pytest
import torch

from oracle import cave_oracle as O

pytestmark = pytest.mark.gpu


class _Model:
    def __init__(self, sense):
        self.modelSense = sense


def _data(kind, B, seed, regime="near"):
    from cave_b200 import synth
    insts = synth.make_batch(kind, B, seed=seed)
    return synth.densify(insts).numpy(), synth.predictions(insts, seed, regime)


def test_get_projection_exact_and_inner_match_oracle_targets():
    from cave_b200 import EPO, exactConeAlignedCosine, innerConeAlignedCosine
    dev = torch.device("cuda:0")
    ctrs, pred = _data("tsp20", 12, 5)
    c = -pred.astype(np.float64)                       # signed cost (MINIMIZE)
    ct, At = torch.tensor(c, device=dev), torch.tensor(ctrs, device=dev)
    t_ref, _, _ = O.exact_target(c, ctrs.astype(np.float64), fp64=True)
    t = exactConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda")._get_projection(ct, At)
    np.testing.assert_allclose(t.cpu().numpy(), t_ref, rtol=1e-5, atol=1e-7)
    # inner, QP branch (solve_ratio = 1: the draw is consumed, the branch is always the projection)
    t_ref, _, rn = O.inner_target(c, ctrs.astype(np.float64), 0.2, fp64=True)
    mod = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", inner_ratio=0.2, solve_ratio=1.0, seed=3)
    np.testing.assert_allclose(mod._get_projection(ct, At).cpu().numpy(), t_ref, rtol=1e-5, atol=1e-7)
    # an instance inside its cone is returned un-pushed (src/cave.py:218-219)
    lam = np.random.default_rng(0).random(ctrs.shape[1])
    c_in = (lam @ ctrs[0].astype(np.float64))[None]
    t_in, _, rn_in = O.inner_target(c_in, ctrs[:1].astype(np.float64), 0.2, fp64=True)
    assert rn_in[0] < 1e-7
    got = mod._get_projection(torch.tensor(c_in, device=dev), At[:1]).cpu().numpy()
    np.testing.assert_allclose(got, t_in, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(np.linalg.norm(got), 1.0, rtol=1e-6)            # normalised projection, no avg mixed in
    # heuristic branch (solve_ratio = 0: uniform() > 0 always)
    modh = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", inner_ratio=0.2, solve_ratio=0.0, seed=3)
    t_h = O.heuristic_target(c, ctrs.astype(np.float64), 0.2)
    np.testing.assert_allclose(modh._get_projection(ct, At).cpu().numpy(), t_h, rtol=1e-5, atol=1e-6)


def test_vrp20_hybrid_sequence_follows_the_seeded_branch_draws():
    """configs[3]: CaVE Hybrid, solve_ratio = 0.3, ragged VRP-20 rows.  One host draw per forward call
    (src/cave.py:201); the oracle replays the same RandomState, so every call must land on the same branch and
    the same loss / gradient."""
    from cave_b200 import EPO, innerConeAlignedCosine
    dev = torch.device("cuda:0")
    seed, ratio = 7, 0.3
    mod = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", inner_ratio=0.2, solve_ratio=ratio, seed=seed,
                                 solver_kwargs={"precision": "fp64"})
    rng = np.random.RandomState(seed)
    branches = []
    for call in range(10):
        ctrs, pred = _data("vrp20", 16, 100 + call, "near" if call % 2 else "uniform")
        mode = O.MODE_HEURISTIC if rng.uniform() > ratio else O.MODE_INNER
        branches.append(mode)
        ref = O.forward_backward(pred.astype(np.float64), ctrs, mode=mode, inner_ratio=0.2, fp64=True)
        p = torch.tensor(pred.astype(np.float64), device=dev, requires_grad=True)
        loss = mod(p, torch.tensor(ctrs, device=dev))
        loss.backward()
        np.testing.assert_allclose(loss.item(), ref["loss"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref["grad"], rtol=1e-5, atol=1e-5 * np.abs(ref["grad"]).max())
    assert O.MODE_HEURISTIC in branches and O.MODE_INNER in branches       # the sequence exercised both


def test_pack_must_belong_to_the_tensor_and_index_is_range_checked():
    """a warm pack reused for another same-shape batch must raise, and an out-of-range dataset index
    must be reported (status + NaN), never read out of bounds."""
    from cave_b200 import _lib, cave_forward_backward, pack_constraints
    dev = torch.device("cuda:0")
    ctrs, pred = _data("tsp20", 8, 1)
    A = torch.tensor(ctrs, device=dev)
    pk = pack_constraints(A)
    p = torch.tensor(pred, device=dev)
    cave_forward_backward(p, A, -1.0, 1, pack=pk)
    with pytest.raises(ValueError):
        cave_forward_backward(p, A.clone(), -1.0, 1, pack=pk)
    A2 = A.clone()
    pk2 = pack_constraints(A2)
    A2[0, 0, 0] = 5.0
    with pytest.raises(ValueError):
        cave_forward_backward(p, A2, -1.0, 1, pack=pk2)
    idx = torch.tensor([0, 3, 8, -1, 7, 2 ** 31 - 1], dtype=torch.int64, device=dev)
    out = cave_forward_backward(p[:6], None, -1.0, 1, 0.2, "none", pack=pk, index=idx, want_status=True)
    st = (out["status"] & 0xff).cpu().tolist()
    assert st[2] == _lib.ST_BADINPUT and st[3] == _lib.ST_BADINPUT and st[5] == _lib.ST_BADINPUT
    assert st[0] == 0 and st[1] == 0 and st[4] == 0
    assert bool(torch.isnan(out["grad"][2]).all()) and bool(torch.isfinite(out["grad"][[0, 1, 4]]).all())


def test_strict_mode_raises_on_unconverged_instances():
    from cave_b200 import EPO, innerConeAlignedCosine
    dev = torch.device("cuda:0")
    ctrs, pred = _data("tsp20", 4, 2)
    mod = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", seed=0, solver_kwargs={"strict": True, "max_iter": 1})
    with pytest.raises(RuntimeError):
        mod(torch.tensor(pred, device=dev), torch.tensor(ctrs, device=dev))
    ok = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", seed=0, solver_kwargs={"strict": True})
    assert torch.isfinite(ok(torch.tensor(pred, device=dev), torch.tensor(ctrs, device=dev)))


def test_two_devices_in_one_process():
    """kernel attributes and the SM count are per device, not per process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cave_b200 import cave_forward_backward
    ctrs, pred = _data("tsp50", 2, 4)
    outs = []
    for i in (0, 1):
        dev = torch.device("cuda", i)
        outs.append(cave_forward_backward(torch.tensor(pred, device=dev), torch.tensor(ctrs, device=dev), -1.0, 1)["grad"].cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("kind,B", [("tsp20", 24), ("vrp20", 24), ("sp5", 16), ("tsp50", 3)])
def test_sparse_ingestion_matches_the_dense_path_bitwise(kind, B):
    """cave_pack_sparse (per-instance CSR in, no dense tensor on the device) must give the pack of the dense scan:
    integer-valued rows => identical results, bit for bit, through the warm / indexed solve."""
    from cave_b200 import SparseConstraints, cave_forward_backward, pack_constraints_sparse, synth
    dev = torch.device("cuda:0")
    insts = synth.make_batch(kind, B, seed=9)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 9, "near"), device=dev, dtype=torch.float64)
    ref = cave_forward_backward(pred, A, -1.0, 1, 0.2, "none", want_proj=True, want_status=True)
    sc = SparseConstraints.from_instances(insts)
    assert sc.nbytes() < A.numel() * 4
    pk = pack_constraints_sparse(sc)
    out = cave_forward_backward(pred, None, -1.0, 1, 0.2, "none", want_proj=True, want_status=True, pack=pk)
    one = cave_forward_backward(pred, sc.pin_memory(), -1.0, 1, 0.2, "none", want_proj=True, want_status=True)    # one-shot from host CSR
    for o in (out, one):
        assert ((o["status"] & 0xff) == 0).all()
        for k in ("loss_i", "grad", "proj", "rnorm"):
            assert torch.equal(o[k], ref[k]), k
    # a permuted index into the same pack
    idx = torch.randperm(B, device=dev).to(torch.int32)
    sub = cave_forward_backward(pred[idx.long()], None, -1.0, 1, 0.2, "none", pack=pk, index=idx)
    assert torch.equal(sub["grad"], ref["grad"][idx.long()])


def test_sparse_ingestion_float_rows_and_instances_that_need_the_dense_rows():
    from cave_b200 import _lib, SparseConstraints, cave_forward_backward, pack_constraints_sparse
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    B, m, d = 6, 40, 30
    A = np.zeros((B, m, d), np.float32)
    for b in range(B):
        for r in range(12):                                     # general rows with float values
            k = rng.choice(d, size=rng.integers(2, 6), replace=False)
            A[b, r, k] = rng.standard_normal(len(k)).astype(np.float32)
        for k in range(d):                                      # one singleton row per coordinate
            A[b, 12 + k if 12 + k < m else m - 1, k] = rng.choice([-1.5, 2.0])
    A[5, 12:] = 0                                               # instance 5: no singleton row -> needs the dense rows
    At = torch.tensor(A, device=dev)
    pred = torch.tensor(rng.standard_normal((B, d)), device=dev)
    ref = cave_forward_backward(pred, At, -1.0, 0, reduction="none", want_proj=True, want_status=True)
    pk = pack_constraints_sparse(SparseConstraints.from_dense(At))
    out = cave_forward_backward(pred, None, -1.0, 0, reduction="none", want_proj=True, want_status=True, pack=pk)
    st = (out["status"] & 0xff).cpu().tolist()
    assert st[:5] == [0] * 5 and st[5] == _lib.ST_NOSPACE
    np.testing.assert_allclose(out["proj"][:5].cpu().numpy(), ref["proj"][:5].cpu().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["loss_i"][:5].cpu().numpy(), ref["loss_i"][:5].cpu().numpy(), rtol=1e-10, atol=1e-12)


def test_module_accepts_sparse_constraints_and_packs():
    from cave_b200 import EPO, SparseConstraints, innerConeAlignedCosine, pack_constraints_sparse, synth
    dev = torch.device("cuda:0")
    insts = synth.make_batch("tsp20", 8, seed=4)
    A = synth.densify(insts, device=dev)
    pred = synth.predictions(insts, 4, "near")
    mod = innerConeAlignedCosine(_Model(EPO.MINIMIZE), solver="cuda", seed=0)
    # recycled allocator memory must not be mistaken for a cached solver setup: leave valid-looking garbage behind
    junk = torch.ones(64 << 20, dtype=torch.int32, device=dev)
    del junk
    p0 = torch.tensor(pred, device=dev, requires_grad=True)
    l0 = mod(p0, A); l0.backward()
    sc = SparseConstraints.from_instances(insts)
    p1 = torch.tensor(pred, requires_grad=True)                 # host prediction, host CSR
    l1 = mod(p1, sc); l1.backward()
    p2 = torch.tensor(pred, device=dev, requires_grad=True)
    l2 = mod(p2, pack_constraints_sparse(sc)); l2.backward()
    assert l1.device.type == "cpu" and float(l1) == float(l0) == float(l2)
    assert torch.equal(p1.grad, p0.grad.cpu()) and torch.equal(p2.grad, p0.grad)


def test_held_karp_is_exact_against_brute_force_and_bounds():
    """cave_tsp_solve: n = 8 against all 5040 tours; n = 20 against 2-opt improved tours (an upper bound) and the
    consistency of objective, tour and edge incidence."""
    import itertools
    from cave_b200 import synth, tsp_exact
    rng = np.random.default_rng(0)
    n = 8
    d = n * (n - 1) // 2
    c = rng.random((5, d)).astype(np.float32) + 0.1
    sol, obj, tours = tsp_exact.solve(c, n)
    eidx = synth._edge_index(n)
    for i in range(5):
        best = min(sum(c[i, eidx[p[k], p[(k + 1) % n]]] for k in range(n)) for p in
                   ((0,) + q for q in itertools.permutations(range(1, n))))
        assert abs(obj[i] - best) < 1e-5
        assert sorted(tours[i].tolist()) == list(range(n)) and sol[i].sum() == n
        assert abs((sol[i] * c[i].astype(np.float64)).sum() - obj[i]) < 1e-9
    n = 20
    x, c = tsp_exact.gen_data(6, 10, n, 4, 0.5, seed=1)
    sol, obj, tours = tsp_exact.solve(c, n)
    eidx = synth._edge_index(n)
    for i in range(6):
        assert sorted(tours[i].tolist()) == list(range(n)) and sol[i].sum() == n
        t = list(rng.permutation(n))                      # 2-opt from a random tour: never better than the optimum
        length = lambda t_: sum(c[i, eidx[t_[k], t_[(k + 1) % n]]] for k in range(n))  # noqa: E731
        improved = True
        while improved:
            improved = False
            for a_ in range(n - 1):
                for b_ in range(a_ + 2, n):
                    t2 = t[:a_ + 1] + t[a_ + 1:b_ + 1][::-1] + t[b_ + 1:]
                    if length(t2) < length(t) - 1e-9:
                        t, improved = t2, True
        assert obj[i] <= length(t) + 1e-6
        assert obj[i] <= length(list(range(n))) + 1e-6


def test_no_write_outside_the_callers_buffers():
    """compute-sanitizer is closed on this pool, so the out-of-bounds check is our own: pack, scratch and every output of the
    C-ABI call sit between poisoned guard bands that must come back untouched — structured (Newton), Lawson-Hanson, the dense
    tensor-core path (two rounds) and a dataset-indexed call."""
    import ctypes
    from cave_b200 import _lib, synth
    lib = _lib.load()
    dev = torch.device("cuda:0")
    G = 1 << 16                                   # guard band, bytes

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * G + 512,), 0xA5, dtype=torch.uint8, device=dev)
        off = (-buf.data_ptr() - G) % 256 + G      # 256-byte aligned payload
        return buf, off

    def check(buf, off, nbytes, what):
        assert bool((buf[:off] == 0xA5).all()) and bool((buf[off + nbytes:] == 0xA5).all()), f"guard band of {what} overwritten"

    def run(A, pred, mode, opts, index=None, n_packed=0):
        B, m, d = A.shape
        Bq = pred.shape[0]
        nb = ctypes.c_size_t()
        _lib.check(lib.cave_pack_bytes(B, m, d, ctypes.byref(nb))); pbytes = nb.value
        _lib.check(lib.cave_scratch_bytes(Bq, m, d, _lib.F64, ctypes.byref(opts), ctypes.byref(nb))); sbytes = nb.value
        bufs = {"pack": guarded(pbytes) + (pbytes,), "scratch": guarded(sbytes) + (sbytes,)}
        for name, n in (("loss", 4), ("loss_i", 4 * Bq), ("grad", 4 * Bq * d), ("proj", 4 * Bq * d), ("rnorm", 4 * Bq), ("status", 4 * Bq), ("iters", 4 * Bq)):
            bufs[name] = guarded(n) + (n,)
        ptr = lambda k: ctypes.c_void_p(bufs[k][0].data_ptr() + bufs[k][1])  # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if index is not None:
            _lib.check(lib.cave_pack(ctypes.c_void_p(A.data_ptr()), None, B, m, d, ptr("pack"), pbytes, stream))
        _lib.check(lib.cave_forward_backward(ctypes.c_void_p(A.data_ptr()), None, ctypes.c_void_p(pred.data_ptr()), Bq, m, d, -1.0, mode, 0.2, 0,
                                             _lib.F32, _lib.F64, ctypes.byref(opts), ptr("loss"), ptr("loss_i"), ptr("grad"), ptr("proj"),
                                             ptr("rnorm"), ptr("status"), ptr("iters"), ptr("pack"), pbytes, ptr("scratch"), sbytes, stream))
        torch.cuda.synchronize()
        for k, (buf, off, n) in bufs.items():
            check(buf, off, n, k)
        st = bufs["status"][0][bufs["status"][1]:bufs["status"][1] + 4 * Bq].view(torch.int32)
        return st.cpu().numpy()

    insts = synth.make_batch("tsp20", 12, seed=2)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 2, "near"), device=dev)
    st = run(A, pred, 1, _lib.SolverOpts(0, 0, 0.0, 0, 0, 0, 0, None, 0, 0))
    assert ((st & 0xff) == 0).all()
    idx = torch.tensor([5, 0, 11, 3], dtype=torch.int32, device=dev)
    st = run(A, pred[idx.long()].contiguous(), 1, _lib.SolverOpts(0, 0, 0.0, 0, 0, 1, 0, idx.data_ptr(), 12, 0), index=idx)
    assert ((st & 0xff) == 0).all()
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.randn((5, 40, 24), generator=g, device=dev); c = torch.randn((5, 24), generator=g, device=dev)
    st = run(A, c, 0, _lib.SolverOpts(0, 0, 0.0, 0, 0, 0, 0, None, 0, 0))
    assert ((st & _lib.ST_PATH_LH) != 0).all()
    A = torch.randn((5, 200, 260), generator=g, device=dev); A[3, 190:] = 0; c = torch.randn((5, 260), generator=g, device=dev)
    st = run(A, c, 0, _lib.SolverOpts(0, 0, 0.0, 0, 0, 0, 1, None, 0, 2))            # dense path, slots = 2 -> three rounds
    assert ((st & _lib.ST_PATH_GRAM) != 0).all() and ((st & 0xff) == 0).all()


def test_one_shot_exact_pack_has_no_average_and_is_refused_by_the_other_modes():
    """A cold cave_forward_backward in CAVE_MODE_EXACT packs without the average unit normal (the exact loss never reads it,
    src/cave.py:84-129).  The pack is marked: reusing it (warm) for the exact mode is fine, for CaVE+ / heuristic every
    instance reports CAVE_ST_BADINPUT with NaN outputs instead of being pushed towards a zero average."""
    import ctypes
    from cave_b200 import _lib, synth
    from cave_b200.qpsolver import _opts, _ptr
    lib = _lib.load()
    dev = torch.device("cuda:0")
    insts = synth.make_batch("tsp20", 5, seed=11)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 11, "near"), device=dev, dtype=torch.float32)
    B, m, d = A.shape
    nb = ctypes.c_size_t()
    _lib.check(lib.cave_pack_bytes(B, m, d, ctypes.byref(nb)))
    pack = torch.empty(nb.value, dtype=torch.uint8, device=dev)

    def call(mode, warm):
        opts = _opts(warm=warm)
        _lib.check(lib.cave_scratch_bytes(B, m, d, 1, ctypes.byref(opts), ctypes.byref(nb)))
        scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        loss_i = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty((B, d), dtype=torch.float32, device=dev)
        rnorm = torch.empty(B, dtype=torch.float32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(lib.cave_forward_backward(
            _ptr(A), None, _ptr(pred), B, m, d, -1.0, mode, 0.2, _lib.REDUCE["mean"], _lib.F32, 1, ctypes.byref(opts),
            _ptr(loss), _ptr(loss_i), _ptr(grad), None, _ptr(rnorm), _ptr(status), _ptr(iters), _ptr(pack), pack.numel(),
            _ptr(scratch), scratch.numel(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.synchronize()
        return status.cpu().numpy() & 0xff, float(loss), grad

    st_cold, loss_cold, grad_cold = call(_lib.MODE_EXACT, warm=False)
    assert (st_cold == _lib.ST_CONVERGED).all()
    st_warm, loss_warm, grad_warm = call(_lib.MODE_EXACT, warm=True)
    assert (st_warm == _lib.ST_CONVERGED).all() and loss_warm == loss_cold and torch.equal(grad_warm, grad_cold)
    st_inner, loss_inner, grad_inner = call(_lib.MODE_INNER, warm=True)
    assert (st_inner == _lib.ST_BADINPUT).all() and np.isnan(loss_inner) and bool(torch.isnan(grad_inner).all())
    # a cold call in the pushing mode writes the full pack again
    st2, loss2, _ = call(_lib.MODE_INNER, warm=False)
    assert (st2 == _lib.ST_CONVERGED).all() and np.isfinite(loss2)
    st3, loss3, _ = call(_lib.MODE_INNER, warm=True)
    assert (st3 == _lib.ST_CONVERGED).all() and loss3 == loss2
