"""Host-side mirror of the reference's module interface (src/cave.py:93-119, 152-195; error
behaviour of test/test_func.py:195-227).  No kernel is launched here."""
import numpy as np
import pytest
import torch

from cave_b200 import EPO, exactConeAlignedCosine, innerConeAlignedCosine
from cave_b200 import _lib


class _Model:
    def __init__(self, sense=EPO.MINIMIZE):
        self.modelSense = sense


def test_invalid_solver_raises():
    with pytest.raises(ValueError):
        exactConeAlignedCosine(_Model(), solver="bogus")            # test/test_func.py:195-200


@pytest.mark.parametrize("solver", ["nnls", "clarabel", "apgd"])
def test_reference_cpu_backends_are_not_silently_provided(solver):
    with pytest.raises(ValueError, match="never falls back"):
        exactConeAlignedCosine(_Model(), solver=solver)


def test_invalid_ratios_raise():
    with pytest.raises(ValueError):
        innerConeAlignedCosine(_Model(), solve_ratio=1.5)           # test/test_func.py:202-207
    with pytest.raises(ValueError):
        innerConeAlignedCosine(_Model(), inner_ratio=-0.1)          # test/test_func.py:209-214
    with pytest.raises(ValueError):
        innerConeAlignedCosine(_Model(), solver_kwargs={"bogus": 1})


def test_constructor_contract_matches_reference():
    m = innerConeAlignedCosine(_Model(), solver="cuda", solver_kwargs=None, max_iter=3, solve_ratio=0.3,
                               inner_ratio=0.25, processes=4, reduction="sum", seed=7)
    assert (m.solver, m.max_iter, m.solve_ratio, m.inner_ratio, m.processes, m.reduction) == \
        ("cuda", 3, 0.3, 0.25, 4, "sum")
    assert m.solver_kwargs == {} and m.pool is None
    e = exactConeAlignedCosine(_Model(), solver_kwargs={"precision": "fp32"})
    assert e.solver == "cuda" and e.solver_kwargs == {"precision": "fp32"}


def test_branch_sequence_follows_randomstate_and_is_consumed_every_call():
    # src/cave.py:201: one uniform() draw per forward call, even when solve_ratio == 1
    a = innerConeAlignedCosine(_Model(), solve_ratio=0.5, seed=7)
    b = innerConeAlignedCosine(_Model(), solve_ratio=0.5, seed=7)
    seq_a = [a._mode() for _ in range(32)]
    assert seq_a == [b._mode() for _ in range(32)]
    rs = np.random.RandomState(7)
    assert seq_a == [_lib.MODE_HEURISTIC if rs.uniform() > 0.5 else _lib.MODE_INNER for _ in range(32)]
    c = innerConeAlignedCosine(_Model(), solve_ratio=1.0, seed=3)
    rs = np.random.RandomState(3)
    for _ in range(5):
        assert c._mode() == _lib.MODE_INNER
        rs.uniform()
    assert c._branch_rng.uniform() == rs.uniform()
    assert innerConeAlignedCosine(_Model(), solve_ratio=0.0, seed=1)._mode() == _lib.MODE_HEURISTIC


def test_invalid_model_sense_raises_at_call():
    m = exactConeAlignedCosine(_Model(sense="sideways"))
    with pytest.raises(ValueError, match="modelSense"):
        m(torch.zeros(2, 3), torch.zeros(2, 4, 3))


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = exactConeAlignedCosine(_Model())
    with pytest.raises(_lib.CaveLibraryError, match="no CPU fallback"):
        m(torch.rand(2, 3), torch.rand(2, 4, 3))


def test_shape_validation():
    from cave_b200 import cave_forward_backward
    with pytest.raises(ValueError, match="shape mismatch"):
        cave_forward_backward(torch.zeros(2, 3), torch.zeros(2, 4, 5), -1.0, 0)
