"""cave_b200 — B200-native backend (solver='cuda') for the CaVE cone-projection hot path.

Public surface (mirrors the reference's src/cave.py and src/qpsolver.py for this path):
    exactConeAlignedCosine, innerConeAlignedCosine      loss modules, solver='cuda'
    project_cuda(tight_ctrs, signed_cost, **kw)         batched projection -> (proj, rnorm)
    cave_forward_backward(...)                          fused loss + gradient, one C-ABI call
    pack_constraints(tight_ctrs)                        device-resident packed constraints
    SparseConstraints / pack_constraints_sparse(sc)     the same pack from per-instance CSR (no dense tensor)
The CUDA library (cave_b200/_C/libcave_b200.so) is loaded on first use; if it is missing the
call raises — there is no CPU fallback.
"""
from ._pyepo_compat import EPO, optModel, optModule  # noqa: F401
from .cave import abstractConeAlignedCosine, exactConeAlignedCosine, innerConeAlignedCosine  # noqa: F401
from .qpsolver import (CavePack, SparseConstraints, cave_forward_backward, pack_constraints,  # noqa: F401
                       pack_constraints_sparse, project_cuda)

__all__ = ["EPO", "optModel", "optModule", "abstractConeAlignedCosine", "exactConeAlignedCosine",
           "innerConeAlignedCosine", "project_cuda", "cave_forward_backward", "pack_constraints", "CavePack",
           "SparseConstraints", "pack_constraints_sparse"]
