"""Parity metrics of SURVEY.md section 8(d), shared by the tests, smoke() and bench.py."""
import numpy as np


def sv_rel_err(d, d_ref):
    d, d_ref = np.asarray(d), np.asarray(d_ref)
    return float(np.max(np.abs(d - d_ref) / d_ref))


def sign_align(A, A_ref):
    """Flip each column of A so that <a_i, a_ref_i> >= 0; returns (A_aligned, flips)."""
    s = np.sign(np.sum(A * A_ref, axis=0))
    s[s == 0] = 1.0
    return A * s[None, :], s


def column_angles(A, A_ref):
    """Per-column angle (rad) after sign alignment, via asin of the residual norm
    (acos of the cosine bottoms out near 1e-8)."""
    A, _ = sign_align(A, A_ref)
    a = A / np.linalg.norm(A, axis=0, keepdims=True)
    b = A_ref / np.linalg.norm(A_ref, axis=0, keepdims=True)
    r = np.linalg.norm(a - b * np.sum(a * b, axis=0, keepdims=True), axis=0)
    return np.arcsin(np.clip(r, 0.0, 1.0))


def subspace_angle(A, A_ref):
    """Largest principal angle (rad) between the column spaces of A and A_ref
    (QR, then the sine from the residual of projecting one basis on the other)."""
    Q1, _ = np.linalg.qr(np.asarray(A, dtype=np.float64))
    Q2, _ = np.linalg.qr(np.asarray(A_ref, dtype=np.float64))
    R = Q1 - Q2 @ (Q2.T @ Q1)
    s = np.linalg.svd(R, compute_uv=False)
    return float(np.arcsin(np.clip(s[0], 0.0, 1.0)))
