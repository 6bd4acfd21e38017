"""
Emulation of the reference's DEFAULT CaVE+ backend, ``solver='clarabel'`` truncated at ``max_iter=3``
(/root/reference/src/cave.py:157, 277-295; README.md:38).  **Parity unpinned**: cvxpy and Clarabel are not installed in
this image and cannot be fetched, so nothing here can be checked against the real solver.  What is emulated is the
SEMANTICS the method relies on: a primal-dual interior-point method for

    min_{lam >= 0}  || c - lam @ A ||^2            (cvxpy: sum_squares(cp_param - lam_var @ ctr_param), cave.py:273)

stopped after three predictor-corrector iterations returns a strictly interior iterate lam > 0 — a point INSIDE the cone
near the projection — which is why the clarabel branch of ``innerConeAlignedCosine._get_projection`` returns the
normalised projection without the extra push-inside step (cave.py:213-215).  Implementation: Mehrotra predictor-corrector
on the KKT system  G lam - b - z = 0, lam o z = mu,  G = A A^T, b = A c, after Jacobi equilibration (Clarabel equilibrates
its KKT matrix), unit starting point, step length 0.99 of the distance to the boundary; batched torch.linalg on the GPU.
It is a comparison arm for examples/train_tsp20_regret.py, not part of the product path.
"""
import torch


def project_ipm_truncated(A: torch.Tensor, c: torch.Tensor, iters: int = 3) -> torch.Tensor:
    """A [B, m, d] (zero rows = padding), c [B, d] -> approximate projection [B, d] after `iters` interior-point steps."""
    A = A.double(); c = c.double()
    B, m, d = A.shape
    valid = (A.abs().sum(dim=2) > 1e-7)
    nrm = A.norm(dim=2).clamp(min=1e-12)
    As = A / nrm[:, :, None]                                  # Jacobi scaling: unit rows
    G = As @ As.transpose(1, 2)
    b = (As @ c[:, :, None]).squeeze(2)
    eye = torch.eye(m, dtype=A.dtype, device=A.device)[None]
    # padding rows: decoupled unit diagonal, zero right-hand side
    G = torch.where((valid[:, :, None] & valid[:, None, :]), G, eye.expand(B, m, m) * 1.0)
    b = torch.where(valid, b, torch.zeros_like(b))
    lam = torch.ones((B, m), dtype=A.dtype, device=A.device)
    z = torch.ones_like(lam)
    for _ in range(iters):
        r_d = (G @ lam[:, :, None]).squeeze(2) - b - z        # dual residual
        mu = (lam * z).sum(1, keepdim=True) / m
        if float(mu.max()) < 1e-13:                          # converged to machine precision (only reached with many iterations)
            break
        H = G + torch.diag_embed(z / lam + 1e-12)
        # predictor (affine scaling)
        rhs = -r_d - z
        dl_a = torch.linalg.solve(H, rhs[:, :, None]).squeeze(2)
        dz_a = -z - z / lam * dl_a
        a_p = _step(lam, dl_a); a_d = _step(z, dz_a)
        mu_a = ((lam + a_p * dl_a) * (z + a_d * dz_a)).sum(1, keepdim=True) / m
        sigma = (mu_a / mu).clamp(max=1.0) ** 3
        # corrector
        comp = sigma * mu - dl_a * dz_a
        rhs = -r_d - z + comp / lam
        dl = torch.linalg.solve(H, rhs[:, :, None]).squeeze(2)
        dz = (comp - z * dl) / lam - z
        a_p = 0.99 * _step(lam, dl); a_d = 0.99 * _step(z, dz)
        lam = lam + a_p * dl
        z = z + a_d * dz
    lam = torch.where(valid, lam / nrm, torch.zeros_like(lam))
    return (lam[:, None, :] @ A).squeeze(1)


def _step(x, dx):
    """Largest step in [0, 1] keeping x + a dx >= 0, per batch row."""
    ratio = torch.where(dx < 0, -x / dx, torch.full_like(x, float("inf")))
    return ratio.min(dim=1, keepdim=True).values.clamp(max=1.0)


class ClarabelTruncatedCosine(torch.nn.Module):
    """innerConeAlignedCosine with the emulated ``solver='clarabel', max_iter=3`` branch: loss = 1 - cos(c, proj / ||proj||)
    (cave.py:72, 213-215); sign convention and reduction as in cave.py:62-73."""

    def __init__(self, minimize: bool = True, max_iter: int = 3, reduction: str = "mean"):
        super().__init__()
        self.sign, self.max_iter, self.reduction = (-1.0 if minimize else 1.0), max_iter, reduction

    def forward(self, pred_cost, tight_ctrs):
        c = self.sign * pred_cost
        with torch.no_grad():
            proj = project_ipm_truncated(tight_ctrs, c, self.max_iter).to(c.dtype)
            t = proj / proj.norm(dim=1, keepdim=True).clamp(min=1e-8)
        loss = 1.0 - torch.nn.functional.cosine_similarity(c, t, dim=1)
        return loss.mean() if self.reduction == "mean" else (loss.sum() if self.reduction == "sum" else loss)
