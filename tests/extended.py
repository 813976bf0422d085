"""Extended-precision (x87 80-bit long double) evaluation of the affine Riccati sweep, used by the tests to
measure how well-determined the gains are: |K_float64 - K_longdouble| is the rounding noise of ANY float64
implementation of trajectory_generation.py:183-216 on a given iterate (reference and GPU alike)."""
import numpy as np

from oracle import acro_oracle as O

LD = np.longdouble


def stage_lists_ld(x, u, x_ref, u_ref, Q, R, Q_T, m=O.DEFAULT):
    """Same closed form as the oracle, evaluated in long double."""
    x, u, x_ref, u_ref = (np.asarray(a, dtype=LD) for a in (x, u, x_ref, u_ref))
    Q, R, Q_T = (np.asarray(a, dtype=LD) for a in (Q, R, Q_T))
    a1, h, a3, g1, g2, f1, f2, dt = (LD(v) for v in (m.a1, m.h, m.a3, m.g1, m.g2, m.f1, m.f2, m.dt))
    xs, us = x[:-1], u
    th1, th2, w1, w2 = xs[:, 0], xs[:, 1], xs[:, 2], xs[:, 3]
    s1, c1, s2, c2 = np.sin(th1), np.cos(th1), np.sin(th2), np.cos(th2)
    s12, c12 = np.sin(th1 + th2), np.cos(th1 + th2)
    M11, M12, M22 = a1 + 2 * h * c2, a3 + h * c2, a3 + 0 * c2
    det = M11 * M22 - M12 * M12
    r1 = h * s2 * w2 * w1 + h * s2 * (w1 + w2) * w2 - f1 * w1 - (g1 * s1 + g2 * s12)
    r2 = us[:, 1] - h * s2 * w1 * w1 - f2 * w2 - g2 * s12
    dd1, dd2 = (M22 * r1 - M12 * r2) / det, (M11 * r2 - M12 * r1) / det
    cols = [(-(g1 * c1 + g2 * c12), -g2 * c12),
            (h * c2 * (w1 * w2 + (w1 + w2) * w2) - g2 * c12 + h * s2 * (2 * dd1 + dd2), -h * c2 * w1 * w1 - g2 * c12 + h * s2 * dd1),
            (2 * h * s2 * w2 - f1, -2 * h * s2 * w1), (2 * h * s2 * (w1 + w2), -f2 + 0 * s2)]
    T = xs.shape[0]
    A = np.zeros((T, 4, 4), dtype=LD)
    A[:, 0, 0] = A[:, 1, 1] = A[:, 2, 2] = A[:, 3, 3] = 1
    A[:, 0, 2] = A[:, 1, 3] = dt
    for j, (d0, d1) in enumerate(cols):
        A[:, 2, j] += dt * (M22 * d0 - M12 * d1) / det
        A[:, 3, j] += dt * (M11 * d1 - M12 * d0) / det
    B = np.zeros((T, 4, 2), dtype=LD)
    B[:, 2, 1] = dt * (-M12 / det)
    B[:, 3, 1] = dt * (M11 / det)
    q = 2 * np.einsum("ij,tj->ti", Q, xs - x_ref[:-1])
    r = 2 * np.einsum("ij,tj->ti", R, us - u_ref)
    qT = 2 * Q_T @ (x[-1] - x_ref[-1])
    return A, B, q, r, 2 * Q_T, qT


def solve2(G, F):
    det = G[0, 0] * G[1, 1] - G[0, 1] * G[1, 0]
    inv = np.array([[G[1, 1], -G[0, 1]], [-G[1, 0], G[0, 0]]], dtype=LD) / det
    return inv @ F


def riccati_ld(x, u, x_ref, u_ref, Q=O.Q_NEWTON, R=O.R_NEWTON, Q_T=O.QT_NEWTON):
    A, B, q, r, P, p = stage_lists_ld(x, u, x_ref, u_ref, Q, R, Q_T)
    Q2, R2 = 2 * np.asarray(Q, dtype=LD), 2 * np.asarray(R, dtype=LD)
    T = A.shape[0]
    K = np.zeros((T, 2, 4), dtype=LD)
    S = np.zeros((T, 2), dtype=LD)
    for t in range(T - 1, -1, -1):
        G = R2 + B[t].T @ P @ B[t]
        F = B[t].T @ P @ A[t]
        g = r[t] + B[t].T @ p
        K[t] = -solve2(G, F)
        S[t] = -solve2(G, g)
        p = q[t] + A[t].T @ p - K[t].T @ G @ S[t]
        P = Q2 + A[t].T @ P @ A[t] - K[t].T @ G @ K[t]
    return K, S
