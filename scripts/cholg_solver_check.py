"""Validation in NumPy of the symmetrised dense M-step solve with LLE (tdlo_taskq.cuh: smode 3): with G = L L^T
(Cholesky, once per registration), (S G + c I) W = B  <=>  (L^T S L + c I) V = L^T B,  T - Y0 = G W = L V,  W = L^-T V,
S = diag(D) + eps H.  Compared with a 50-digit dense solve (mpmath) and with LAPACK's dense solve of the unsymmetric system.
Output committed as profiles/r2_cholg_solver_accuracy.txt."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mpmath as mp
import oracle
from trackdlo_b200 import synth
mp.mp.dps = 50

def G_matrix(s, beta):
    d = np.abs(s[:, None] - s[None, :])
    return 1 / (4 * beta * beta) * np.exp(-np.sqrt(2) * d / beta) * (2 * d + np.sqrt(2) * beta)

def cholg_solve(G, S, c, B):
    L = np.linalg.cholesky(G)
    M = L.T @ S @ L + c * np.eye(len(G))
    V = np.linalg.solve(M, L.T @ B)            # SPD: any stable method; the kernel uses unpivoted elimination
    # the kernel's unpivoted Gauss-Jordan, to make sure no pivoting is needed
    n = len(G); AB = np.hstack([M, L.T @ B]).copy()
    for k in range(n):
        piv = AB[k, k]
        for i in range(n):
            if i != k: AB[i] -= AB[k] * (AB[i, k] / piv)
    V2 = AB[:, n:] / np.diag(AB)[:, None]
    import scipy.linalg as sl
    W = sl.solve_triangular(L.T, V2, lower=False)
    return W, L @ V2, np.abs(V - V2).max() / np.abs(V).max()

def mp_ref(G_fn, s, beta, S, c, B):
    n = len(s)
    G = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            dd = abs(mp.mpf(s[i]) - mp.mpf(s[j]))
            G[i, j] = mp.mpf(1) / (4 * mp.mpf(beta) ** 2) * mp.e ** (-mp.sqrt(2) * dd / mp.mpf(beta)) * (2 * dd + mp.sqrt(2) * mp.mpf(beta))
    Sm = mp.matrix(S.tolist())
    A = Sm * G
    for i in range(n): A[i, i] += mp.mpf(c)
    W = mp.matrix(n, B.shape[1])
    for c_ in range(B.shape[1]):
        w = mp.lu_solve(A, mp.matrix(B[:, c_].tolist()))
        for i in range(n): W[i, c_] = w[i]
    V = G * W
    f = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
    return f(W), f(V)

if __name__ == "__main__":
    rng = np.random.default_rng(3)
    print("relative errors (max-norm) against a 50-digit solve: W and G W of the symmetrised solve | of numpy.linalg.solve on the unsymmetric A")
    for (n, beta, sigma2, lam, gamma, occl) in [(50, 3.0, 1e-4, 1.0, 10.0, False), (50, 3.0, 1e-5, 1.0, 10.0, True), (50, 3.0, 2e-3, 1.0, 10.0, True),
                                                (30, 3.0, 1e-4, 1.0, 10.0, False), (64, 3.0, 3e-6, 1.0, 10.0, True), (45, 3.0, 1e-5, 1.0, 10.0, True),
                                                (50, 0.35, 1e-5, 50000.0, 10.0, True), (50, 10.0, 1e-4, 1.0, 1.0, True), (20, 3.0, 1e-7, 1.0, 100.0, False)]:
        f = synth.make_frame(int(rng.integers(0, 1000)), n_nodes=n, n_points=2000)
        Y0 = f["Y"]; s = f["rest"]
        H = oracle.lle_H(Y0)
        D = rng.uniform(0, 2000.0 / n * 20, n)
        if occl: D[n // 3: n // 3 + n // 6] = 0.0; D[2] = 1e-14
        eps = sigma2 * gamma; c = lam * sigma2
        S = np.diag(D) + eps * H
        B = rng.normal(size=(n, 3)) * 0.01 * np.sqrt(D + 1)[:, None] - eps * (H @ Y0)
        G = G_matrix(s, beta)
        try:
            W, V, gjdiff = cholg_solve(G, S, c, B)
        except np.linalg.LinAlgError as e:
            print(f"n={n} beta={beta}: Cholesky of G failed ({e}) -> the kernel falls back to the pivoted path"); continue
        Wr, Vr = mp_ref(G_matrix, s, beta, S, c, B)
        A = S @ G + c * np.eye(n)
        Wn = np.linalg.solve(A, B)
        rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
        print(f"n={n:2d} beta={beta:5.2f} sigma2={sigma2:g} cond(A)={np.linalg.cond(A):.1e} cond(G)={np.linalg.cond(G):.1e}: W {rel(W, Wr):.1e} GW {rel(V, Vr):.1e} | W {rel(Wn, Wr):.1e} GW {rel(G @ Wn, Vr):.1e}   (unpivoted vs LAPACK V: {gjdiff:.1e})")
