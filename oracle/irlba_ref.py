"""CPU restatement of irlba's fast C path (irlba >= 2.3.3, src/irlb.c + R/irlba.R).

TEST INFRASTRUCTURE ONLY (oracle/): the parity checker and the bench's CPU
baseline.  The product package never imports it.

The reference calls `irlba::irlba(A=t(FM), nv=n, center=, scale=)` at
R/utils.R:417 (from `sparse_prcomp_irlba`, R/utils.R:365-430).  irlba is an
un-vendored CRAN dependency (DESCRIPTION:72, `irlba (>= 2.3.3)`), so its
published algorithm (Baglama & Reichel implicitly restarted Lanczos
bidiagonalisation, as coded in irlb.c) is restated here; SURVEY.md Appendix A
is the spec.  Parity is pinned through the reference's own golden values
(tests/testthat/test-preprocessing.R:11,22,34,47-51,69,75,82; test-io.R:267;
test-projection.R:28-30) -- see tests/test_oracle_golden.py.

Matrix naming follows irlba: A is m x n = cells x genes.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


class IrlbaResult(dict):
    __getattr__ = dict.__getitem__


def _orthog(X, y):
    """irlb.c `orthog`: ONE classical Gram-Schmidt pass, y -= X (X^T y)."""
    if X.shape[1] == 0:
        return y
    t = X.T @ y
    return y - X @ t


def _convtests(Bsz, n, tol, svtol, Smax, svratio, res, k, S):
    """irlb.c `convtests`; returns (k, converged)."""
    Len = int(np.sum((np.abs(res[:Bsz]) < tol * Smax) & (svratio[:Bsz] < svtol)))
    if Len >= n or S == 0:
        return k, True
    k = max(k, n + Len)
    k = min(k, Bsz - 3)
    k = max(k, 1)
    return k, False


def irlba(A, nv, v0, center=None, scale=None, work=None, maxit=1000, tol=1e-5,
          svtol=None, matvec_hook=None, rng=None):
    """Top-`nv` singular triplets of (A - 1 center^T) diag(1/scale).

    A: scipy sparse or ndarray, m x n.  v0: start vector of length n (R draws it
    with rnorm(n) after set.seed(2016)).  Returns d[nv], u[m,nv], v[n,nv],
    iter, mprod, err (0 converged, -2 not converged).
    matvec_hook(kind, vec) optionally perturbs products (sensitivity studies).
    """
    m, n = A.shape
    nu = nv
    if svtol is None:
        svtol = tol
    if work is None:
        work = nv + 7
    if nv >= min(m, n):
        raise ValueError("max(nu, nv) must be strictly less than min(nrow(A), ncol(A))")
    if min(m, n) < 6:
        # irlba R/irlba.R: "Tiny problem detected, using standard `svd` function" 
        D = A.toarray() if hasattr(A, "toarray") else np.asarray(A, dtype=np.float64)
        if center is not None:
            D = D - center[None, :]
        if scale is not None:
            D = D / scale[None, :]
        U, S, Vt = np.linalg.svd(D, full_matrices=False)
        return IrlbaResult(d=S[:nu].copy(), u=U[:, :nu].copy(), v=Vt.T[:, :nv].copy(), iter=0, mprod=0, err=0)
    if work <= nv:
        work = nv + 1
    work = min(work, min(m, n))
    if work < 4 or n < 4 or m < 4:
        return IrlbaResult(d=None, u=None, v=None, iter=0, mprod=0, err=-1)
    AT = A.T.tocsr() if hasattr(A, "tocsr") else A.T
    A_ = A.tocsr() if hasattr(A, "tocsr") else A

    def Ax(x):
        xs = x / scale if scale is not None else x
        w = np.asarray(A_ @ xs).ravel()
        if center is not None:
            w = w - np.dot(xs, center)
        if matvec_hook is not None:
            w = matvec_hook("n", w)
        return w

    def Atw(w):
        f = np.asarray(AT @ w).ravel()
        if scale is not None:
            f = f / scale
        if center is not None:
            beta = np.sum(w)
            if scale is not None:
                f = f - beta * center / scale
            else:
                f = f - beta * center
        if matvec_hook is not None:
            f = matvec_hook("t", f)
        return f

    V = np.zeros((n, work))
    W = np.zeros((m, work))
    B = np.zeros((work, work))
    F = np.zeros(n)
    svratio = np.zeros(work)
    k = 0  # "restart" argument of irlb(); 0 on a fresh start
    mprod = 0
    it = 0
    converged = False
    d = np.linalg.norm(v0)
    if d < EPS:
        return IrlbaResult(d=None, u=None, v=None, iter=0, mprod=0, err=-1)
    V[:, 0] = v0 / d
    BU = BS = BVt = None
    S = 0.0
    while it < maxit:
        j = 0 if it == 0 else k
        W[:, j] = Ax(V[:, j])
        mprod += 1
        if it > 0:
            W[:, j] = _orthog(W[:, :j], W[:, j])
        S = np.linalg.norm(W[:, j])
        if S < EPS and j == 0:
            return IrlbaResult(d=None, u=None, v=None, iter=it, mprod=mprod, err=-4)
        W[:, j] *= 1.0 / S
        while j < work:
            F = Atw(W[:, j])
            mprod += 1
            F = F - S * V[:, j]
            F = _orthog(V[:, :j + 1], F)
            if j + 1 < work:
                R_F = np.linalg.norm(F)
                if R_F < EPS:  # near-invariant subspace: fresh random direction
                    if rng is None:
                        raise RuntimeError("invariant subspace hit and no rng supplied")
                    F = _orthog(V[:, :j + 1], rng.rnorm(n))
                    R = 1.0 / np.linalg.norm(F)
                    R_F = 0.0
                else:
                    R = 1.0 / R_F
                V[:, j + 1] = F * R
                B[j, j] = S
                B[j, j + 1] = R_F
                W[:, j + 1] = Ax(V[:, j + 1])
                mprod += 1
                W[:, j + 1] -= R_F * W[:, j]
                W[:, j + 1] = _orthog(W[:, :j + 1], W[:, j + 1])
                S = np.linalg.norm(W[:, j + 1])
                if S < EPS:
                    if rng is None:
                        raise RuntimeError("invariant subspace hit and no rng supplied")
                    wj = _orthog(W[:, :j + 1], rng.rnorm(m))
                    W[:, j + 1] = wj / np.linalg.norm(wj)
                    S = 0.0
                else:
                    W[:, j + 1] *= 1.0 / S
            else:
                B[j, j] = S
            j += 1
        # dgesdd on B (work x work)
        BU, BS, BVt = np.linalg.svd(B)
        R_F = np.linalg.norm(F)
        F = F * (1.0 / R_F)
        if R_F < EPS:
            R_F = 0.0
        Smax = BS.max()
        svratio = np.abs(svratio - BS) / BS
        res = R_F * BU[work - 1, :]
        k, converged = _convtests(work, nu, tol, svtol, Smax, svratio, res, k, S)
        if converged:
            it += 1
            break
        svratio = BS.copy()
        V[:, :k] = V @ BVt.T[:, :k]
        V[:, k] = F
        B[:] = 0.0
        for jj in range(k):
            B[jj, jj] = BS[jj]
            B[jj, k] = res[jj]
        W[:, :k] = W @ BU[:, :k]
        it += 1
    dvals = BS[:nu].copy()
    U = W @ BU[:, :nu]
    Vout = V @ BVt.T[:, :nu]
    return IrlbaResult(d=dvals, u=U, v=Vout, iter=it, mprod=mprod,
                       err=0 if converged else -2)
