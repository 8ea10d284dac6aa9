"""Validation in NumPy of the banded information-form M-step solve with LLE (tdlo_common.cuh: mct_banded_lle_solve, the same
operations in the same order): with z = (f_0, f'_0, f_1, f'_1, ...) the joint values of the Matern-3/2 process and its
derivative at the nodes, K their (block-tridiagonal) precision and P the selection of the f components,
    (S G + c I) W = B,  V = G W      <=>      (c K + P^T S P) z = P^T B,   V = P z,   W = P K z,        S = diag(D) + eps H,
an SPD system of size 2 Nn and half-bandwidth 12 (H = E^T E reaches 6 nodes) -- block LDL^T (one node per pivot) without pivoting, O(Nn).
Compared with a 50-digit dense solve (mpmath) and with LAPACK's dense solve of the unsymmetric A.
Output committed as profiles/r2_banded_solver_accuracy.txt."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, mpmath as mp
import oracle
from trackdlo_b200 import synth
from cholg_solver_check import G_matrix, mp_ref
mp.mp.dps = 50
BW = 12


def q00_factor(x):
    """1 - exp(-2x) (1 + 2x + 2x^2), without the cancellation for small x: exp(-2x) * sum_{k>=3} (2x)^k / k!"""
    if x >= 0.25:
        return 1.0 - np.exp(-2 * x) * (1 + 2 * x + 2 * x * x)
    y = 2 * x; term = y * y * y / 6.0; s = term
    for k in range(4, 16):
        term *= y / k; s += term
    return np.exp(-y) * s


def precision_band(s, beta):
    """K in band storage Kb[i][k] = K[i][i - BW + k] (lower part incl. the diagonal at k = BW), and the f rows of K (6 wide)."""
    n = len(s); a = np.sqrt(2) / beta; s2f = np.sqrt(2) / (4 * beta)
    Kb = np.zeros((2 * n, BW + 1)); Kf = np.zeros((n, 6))          # Kf[t] = K[2t][2t-2 .. 2t+3]
    Kfull = np.zeros((2 * n, 2 * n))
    Kfull[0, 0] = 1 / s2f; Kfull[1, 1] = 1 / (a * a * s2f)
    for t in range(n - 1):
        h = abs(s[t + 1] - s[t]); x = a * h; e = np.exp(-x); e2 = e * e
        Phi = e * np.array([[1 + x, h], [-a * a * h, 1 - x]])
        q00 = s2f * q00_factor(x); q01 = s2f * 2 * a * x * x * e2; q11 = s2f * a * a * (1 - e2 * (1 - 2 * x + 2 * x * x))
        det = q00 * q11 - q01 * q01
        Qi = np.array([[q11, -q01], [-q01, q00]]) / det
        C = Qi @ Phi
        Kfull[2 * t:2 * t + 2, 2 * t:2 * t + 2] += Phi.T @ C
        Kfull[2 * t + 2:2 * t + 4, 2 * t + 2:2 * t + 4] += Qi
        Kfull[2 * t + 2:2 * t + 4, 2 * t:2 * t + 2] = -C
        Kfull[2 * t:2 * t + 2, 2 * t + 2:2 * t + 4] = -C.T
    for i in range(2 * n):
        for k in range(BW + 1):
            j = i - BW + k
            if j >= 0: Kb[i, k] = Kfull[i, j]
    for t in range(n):
        for k in range(6):
            j = 2 * t - 2 + k
            if 0 <= j < 2 * n: Kf[t, k] = Kfull[2 * t, j]
    return Kb, Kf


def banded_solve(s, beta, D, H, eps, c, B):
    """The kernel's algorithm: block LDL^T (2 x 2 pivots = one node per step) right-looking on the band with the right-hand sides
    carried along, multipliers [u v] P^-1 and P^-1 stored in place, then the column-oriented block back substitution."""
    n = len(s); m = 2 * n
    Kb, Kf = precision_band(s, beta)
    A = np.zeros((m + 16, 16)); A[:m, :BW + 1] = c * Kb
    for t in range(n):
        A[2 * t, BW] += D[t]
        for u in range(max(0, t - 6), t + 1):
            A[2 * t, BW - 2 * (t - u)] += eps * H[t, u]
    A[0:m:2, 13:16] = B
    ok = True
    for t in range(n):
        p0, p1 = 2 * t, 2 * t + 1
        a, b, d = A[p0, BW], A[p1, BW - 1], A[p1, BW]
        det = a * d - b * b; inv = 1.0 / det
        rows = list(range(p1 + 1, min(m, p0 + 13)))               # p0+2 .. p0+12
        u = {i: A[i, BW - (i - p0)] for i in rows}; v = {i: A[i, BW - (i - p1)] for i in rows}
        y0 = A[p0, 13:16].copy(); y1 = A[p1, 13:16].copy()
        for i in rows:
            for j in rows:
                if j <= i:
                    A[i, BW - (i - j)] -= (u[i] * (d * u[j] - b * v[j]) + v[i] * (a * v[j] - b * u[j])) * inv
            A[i, 13:16] -= (u[i] * (d * y0 - b * y1) + v[i] * (a * y1 - b * y0)) * inv
        for i in rows:                                            # multipliers [u v] P^-1
            A[i, BW - (i - p0)] = (u[i] * d - v[i] * b) * inv; A[i, BW - (i - p1)] = (v[i] * a - u[i] * b) * inv
        A[p0, BW] = d * inv; A[p1, BW - 1] = -b * inv; A[p1, BW] = a * inv       # P^-1
        ok &= bool(A[p0, BW] > 0 and A[p1, BW] > 0)
    z = np.zeros((m, B.shape[1])); acc = np.zeros_like(z)
    for t in range(n - 1, -1, -1):
        p0, p1 = 2 * t, 2 * t + 1
        r0, r1 = A[p0, 13:16], A[p1, 13:16]
        z[p0] = A[p0, BW] * r0 + A[p1, BW - 1] * r1 - acc[p0]; z[p1] = A[p1, BW - 1] * r0 + A[p1, BW] * r1 - acc[p1]
        for row in (p1, p0):
            for j in range(max(0, row - BW), p0):                 # columns of earlier nodes only
                acc[j] += A[row, BW - (row - j)] * z[row]
    V = z[0::2]
    W = np.zeros_like(V)
    for t in range(n):
        for k in range(6):
            j = 2 * t - 2 + k
            if 0 <= j < m: W[t] += Kf[t, k] * z[j]
    return W, V, ok


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    print("relative errors (max-norm) against a 50-digit solve: W and G W of the banded information-form solve | of numpy.linalg.solve on the unsymmetric A")
    for (n, beta, sigma2, lam, gamma, occl) in [(50, 3.0, 1e-4, 1.0, 10.0, False), (50, 3.0, 1e-5, 1.0, 10.0, True), (50, 3.0, 2e-3, 1.0, 10.0, True),
                                                (30, 3.0, 1e-4, 1.0, 10.0, False), (64, 3.0, 3e-6, 1.0, 10.0, True), (45, 3.0, 1e-5, 1.0, 10.0, True),
                                                (50, 0.35, 1e-5, 50000.0, 10.0, True), (50, 10.0, 1e-4, 1.0, 1.0, True), (20, 3.0, 1e-7, 1.0, 100.0, False),
                                                (200, 3.0, 1e-5, 1.0, 10.0, True), (120, 0.5, 1e-6, 100.0, 10.0, True), (8, 3.0, 1e-4, 1.0, 10.0, False)]:
        f = synth.make_frame(int(rng.integers(0, 1000)), n_nodes=n, n_points=2000)
        Y0 = f["Y"]; s = f["rest"]
        H = oracle.lle_H(Y0)
        D = rng.uniform(0, 2000.0 / n * 20, n)
        if occl and n >= 12: D[n // 3: n // 3 + n // 6] = 0.0; D[2] = 1e-14
        eps = sigma2 * gamma; c = lam * sigma2
        B = rng.normal(size=(n, 3)) * 0.01 * np.sqrt(D + 1)[:, None] - eps * (H @ Y0)
        W, V, ok = banded_solve(s, beta, D, H, eps, c, B)
        S = np.diag(D) + eps * H
        Wr, Vr = mp_ref(G_matrix, s, beta, S, c, B)
        G = G_matrix(s, beta); A = S @ G + c * np.eye(n)
        Wn = np.linalg.solve(A, B)
        rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
        print(f"n={n:3d} beta={beta:5.2f} sigma2={sigma2:g} cond(A)={np.linalg.cond(A):.1e}: banded W {rel(W, Wr):.1e} GW {rel(V, Vr):.1e} pivots>0 {ok} | LAPACK W {rel(Wn, Wr):.1e} GW {rel(G @ Wn, Vr):.1e}")
