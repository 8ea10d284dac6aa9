"""Independent NumPy/LAPACK restatement of cpd_lle (trackdlo.cpp:161-441).

TEST INFRASTRUCTURE ONLY.  Written separately from trackdlo_oracle.cpp (vectorised, LAPACK dgelsy
via scipy as the stand-in for Eigen's COD at trackdlo.cpp:415) so that the two restatements
check each other; the reference has no goldens of its own (parity unpinned).
"""
import numpy as np
import scipy.linalg


def kernel_G(s, beta):
    d = np.abs(s[:, None] - s[None, :])
    return 1.0 / (2 * beta * 2 * beta) * np.exp(-np.sqrt(2.0) * d / beta) * (2 * d + np.sqrt(2.0) * beta)


def cpd_lle(X0, Y, sigma2, beta, lam, lle_weight, mu, max_iter, tol, include_lle=False, priors=None,
            alpha=0.0, vis=None, k_vis=0.0, tau=0.01, H=None, trace=None):
    X0 = np.asarray(X0, float); Y = np.array(Y, float)
    Nn = Y.shape[0]
    d0 = np.sqrt(((Y[:, None, :] - X0[None, :, :]) ** 2).sum(-1))
    X = X0[d0.min(axis=0) < 0.1]
    Mp = X.shape[0]
    Y0 = Y.copy()
    s = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(Y0, axis=0), axis=1))])
    G = kernel_G(s, beta)
    J = np.zeros(Nn); Yext = Y0.copy()
    have_priors = priors is not None and len(priors) > 0
    if have_priors:
        for row in priors:
            i = int(row[0]); J[i] = 1.0; Yext[i] = row[1:4]
    if sigma2 == 0:
        sigma2 = ((Y0[:, None, :] - X[None, :, :]) ** 2).sum() / (3.0 * Nn * Mp)
    use_vis = vis is not None and len(vis) != Nn and len(vis) > 0 and k_vis != 0
    converged = True; W = np.zeros((Nn, 3)); it_done = 0
    cols = np.arange(Mp)
    for it in range(max_iter):
        it_done = it + 1
        diff = Y[:, None, :] - X[None, :, :]
        D2 = (diff ** 2).sum(-1)
        dmin = np.sqrt(D2).min(axis=1)
        dmin[dmin <= tau] = 0.0
        P = np.exp(-0.5 * D2 / sigma2)
        c = (2 * np.pi * sigma2) ** 1.5 * mu / (1 - mu) * Nn / Mp
        P = P / (P.sum(axis=0) + c)
        a = P.argmax(axis=0)
        q1 = np.where(a - 1 == -1, 2, a - 1)
        q2 = np.where(a + 1 == Nn, Nn - 3, a + 1)
        De = np.sqrt(D2)
        b = np.where(De[q1, cols] < De[q2, cols], q1, q2)
        lo = np.minimum(a, b); hi = np.maximum(a, b)
        j = np.arange(Nn)[:, None]
        geo = np.zeros((Nn, Mp))
        below = (np.abs(s[:, None] - s[lo][None, :]) + De[lo, cols][None, :]) ** 2
        above = (np.abs(s[:, None] - s[hi][None, :]) + De[hi, cols][None, :]) ** 2
        geo = np.where(j < lo[None, :], below, geo)
        geo = np.where(j >= hi[None, :], above, geo)
        geo[lo, cols] = D2[lo, cols]
        P = np.exp(-0.5 * geo / sigma2)
        if use_vis:
            v = np.exp(-k_vis * dmin); v = v / v.sum()
            P = P * v[:, None]
            c = (2 * np.pi * sigma2) ** 1.5 * mu / (1 - mu) / Mp
        P = P / (P.sum(axis=0) + c)
        Pt1 = P.sum(axis=0); P1 = P.sum(axis=1); Np = P1.sum(); PX = P @ X
        A = P1[:, None] * G + lam * sigma2 * np.eye(Nn)
        B = PX - P1[:, None] * Y0
        if include_lle:
            A = A + sigma2 * lle_weight * (H @ G)
            B = B - sigma2 * lle_weight * (H @ Y0)
        if have_priors:
            A = A + alpha * J[:, None] * G
            B = B + alpha * (Yext - Y0)
        W = scipy.linalg.lstsq(A, B, lapack_driver="gelsy")[0]
        T = Y0 + G @ W
        sigma2 = ((Pt1 * (X ** 2).sum(1)).sum() - 2 * (PX * T).sum() + (P1 * (T ** 2).sum(1)).sum()) / (Np * 3)
        moved = np.linalg.norm(Y - T, axis=1).sum() / Nn
        Y = T
        if trace is not None:
            trace.append(dict(P1=P1, PX=PX, Np=Np, sigma2=sigma2, W=W.copy(), Y=Y.copy(), A=A, B=B))
        if moved < tol:
            break
        if it == max_iter - 1:
            converged = False
    return dict(Y=Y, sigma2=sigma2, W=W, iters=it_done, converged=converged, kept=Mp)
