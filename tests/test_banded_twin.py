"""The NumPy twin of the kernel's banded information-form M-step solve (scripts/banded_solver_check.py: same operations in the
same order as trackdlo_b200/csrc/tdlo_common.cuh mct_banded_lle_solve) against a dense solve of the reference's system
(S G + c I) W = B, S = diag(D) + eps H (trackdlo.cpp:392-415) -- keeps the derivation under test on the CPU."""
import os
import sys

import numpy as np
import pytest

import oracle
from trackdlo_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
mp = pytest.importorskip("mpmath")            # the script imports it for its 50-digit reference
import banded_solver_check as bsc  # noqa: E402


@pytest.mark.parametrize("n,beta,sigma2,lam,gamma,occl", [(50, 3.0, 1e-4, 1.0, 10.0, False), (45, 3.0, 1e-5, 1.0, 10.0, True),
                                                         (30, 0.35, 1e-5, 50000.0, 10.0, True), (7, 3.0, 1e-4, 1.0, 10.0, False),
                                                         (100, 3.0, 1e-5, 1.0, 10.0, True)])
def test_banded_information_form_equals_dense_solve(n, beta, sigma2, lam, gamma, occl):
    rng = np.random.default_rng(n)
    f = synth.make_frame(n, n_nodes=n, n_points=1500)
    Y0, s = f["Y"], f["rest"]
    H = oracle.lle_H(Y0)
    D = rng.uniform(0, 400.0, n)
    if occl and n >= 12:
        D[n // 3: n // 3 + n // 6] = 0.0
    eps, c = sigma2 * gamma, lam * sigma2
    B = rng.normal(size=(n, 3)) * 0.01 * np.sqrt(D + 1)[:, None] - eps * (H @ Y0)
    W, V, ok = bsc.banded_solve(s, beta, D, H, eps, c, B)
    assert ok
    G = bsc.G_matrix(s, beta)
    A = (np.diag(D) + eps * H) @ G + c * np.eye(n)
    Wd = np.linalg.solve(A, B)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    # the dense solve of the ill-conditioned A is the less accurate of the two (profiles/r2_banded_solver_accuracy.txt)
    assert rel(W, Wd) < 1e-6 and rel(V, G @ Wd) < 1e-6
    # backward error of the twin's solution in the reference's own system
    assert np.abs(A @ W - B).max() <= 1e-12 * (np.abs(A) @ np.abs(W) + np.abs(B)).max()


def test_precision_matrix_is_the_inverse_of_the_kernel_matrix():
    """K = Cov(f_0, f'_0, f_1, ...)^-1: its inverse restricted to the f components is G (trackdlo.cpp:225-233)."""
    s = np.cumsum(np.concatenate([[0.0], np.random.default_rng(0).uniform(0.01, 0.03, 19)]))
    for beta in (0.35, 3.0):
        Kb, _ = bsc.precision_band(s, beta)
        m = 2 * len(s)
        K = np.zeros((m, m))
        for i in range(m):
            for k in range(bsc.BW + 1):
                j = i - bsc.BW + k
                if j >= 0:
                    K[i, j] = K[j, i] = Kb[i, k]
        Gj = np.linalg.inv(K)[0::2, 0::2]
        G = bsc.G_matrix(s, beta)
        assert np.abs(Gj - G).max() / np.abs(G).max() < 1e-7
