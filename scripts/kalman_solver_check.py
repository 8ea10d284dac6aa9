"""Validation of the structured O(Nn) M-step solve (trackdlo_b200/csrc/tdlo_common.cuh: mct_kalman_solve) in NumPy:
the same state-space recursion, against a 50-digit dense solve (mpmath) of (diag(D) G + c I) W = B and against
LAPACK's double-precision dense solve.  Output committed as profiles/r2_kalman_solver_accuracy.txt."""
import numpy as np, mpmath as mp
mp.mp.dps = 50

def G_matrix(s, beta):
    d = np.abs(s[:,None]-s[None,:])
    return 1/(4*beta*beta)*np.exp(-np.sqrt(2)*d/beta)*(2*d+np.sqrt(2)*beta)

def kalman_solve(s, beta, D, c, B):
    """Solves (diag(D) G + c I) W = B with the Matern-3/2 state-space model; returns W and V = G W."""
    n = len(s); a = np.sqrt(2)/beta; sig2 = np.sqrt(2)/(4*beta)
    d = np.sqrt(D)
    bt = np.where(d[:,None] > 0, B/np.where(d>0,d,1)[:,None], 0.0)
    h = np.diff(s)
    Pinf = np.array([[sig2,0],[0,a*a*sig2]])
    Delta = np.zeros((2,2))
    K = np.zeros((n,2)); F = np.zeros(n); Pst = np.zeros((n,2,2)); Phi = np.zeros((n,2,2))
    for t in range(n):
        P = Pinf - Delta
        Pst[t] = P
        F[t] = d[t]*d[t]*P[0,0] + c
        if t < n-1:
            e = np.exp(-a*h[t]); ph = e*np.array([[1+a*h[t], h[t]],[-a*a*h[t], 1-a*h[t]]])
            Phi[t] = ph
            Kp = P[:,0]*d[t]/F[t]
            K[t] = ph @ Kp
            Delta = ph @ Delta @ ph.T + np.outer(K[t],K[t])*F[t]
    # means
    nr = B.shape[1]
    v = np.zeros((n,nr)); am = np.zeros((n,2,nr))
    acur = np.zeros((2,nr))
    for t in range(n):
        am[t] = acur
        v[t] = bt[t] - d[t]*acur[0]
        if t < n-1:
            acur = Phi[t] @ acur + np.outer(K[t], v[t])
    u = np.zeros((n,nr)); V = np.zeros((n,nr))
    r = np.zeros((2,nr))
    for t in range(n-1,-1,-1):
        if t < n-1:
            u[t] = v[t]/F[t] - K[t] @ r
            r = np.vstack([d[t]*u[t], np.zeros(nr)]) + Phi[t].T @ r
        else:
            u[t] = v[t]/F[t]
            r = np.vstack([d[t]*u[t], np.zeros(nr)])
        # r is now r_{t-1}
        V[t] = am[t][0] + Pst[t][0,:] @ r
    W = d[:,None]*u
    return W, V

def dense_ref(s, beta, D, c, B):
    n=len(s)
    G = mp.matrix(n,n)
    for i in range(n):
        for j in range(n):
            dd=abs(mp.mpf(s[i])-mp.mpf(s[j])); G[i,j]= mp.mpf(1)/(4*mp.mpf(beta)**2)*mp.e**(-mp.sqrt(2)*dd/mp.mpf(beta))*(2*dd+mp.sqrt(2)*mp.mpf(beta))
    A = mp.matrix(n,n)
    for i in range(n):
        for j in range(n):
            A[i,j] = mp.mpf(D[i])*G[i,j] + (mp.mpf(c) if i==j else 0)
    W = mp.matrix(n, B.shape[1])
    for c_ in range(B.shape[1]):
        w = mp.lu_solve(A, mp.matrix(B[:,c_].tolist()))
        for i in range(n): W[i,c_] = w[i]
    V = G*W
    f = lambda M: np.array([[float(M[i,j]) for j in range(M.cols)] for i in range(M.rows)])
    return f(W), f(V)

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for (n, beta, L, c, zeros) in [(50,0.35,0.8,0.5,False),(50,0.35,0.8,50000*1e-5,True),(50,3.0,0.8,1e-4,False),(60,0.35,0.8,50000*3e-6,True),(50,0.05,3.0,1e-3,True),(40,0.35,0.8,1e-9,False),(30,10.0,0.5,1e-6,True)]:
        s = np.cumsum(np.concatenate([[0], rng.uniform(0.5,1.5,n-1)])); s *= L/s[-1]
        D = rng.uniform(0, 800, n) * (rng.random(n) < 0.9)
        if zeros: D[10:18] = 0.0; D[3] = 1e-12
        B = rng.normal(size=(n,3)) * np.sqrt(D)[:,None] * 0.01 * 20
        W, V = kalman_solve(s, beta, D, c, B)
        Wr, Vr = dense_ref(s, beta, D, c, B)
        G = G_matrix(s,beta); A = D[:,None]*G + c*np.eye(n)
        Wn = np.linalg.solve(A, B)
        print(f"n={n} beta={beta} c={c:g} cond(A)={np.linalg.cond(A):.2e}: relerr W kalman {np.abs(W-Wr).max()/np.abs(Wr).max():.2e}  numpy-dense {np.abs(Wn-Wr).max()/np.abs(Wr).max():.2e}   V kalman {np.abs(V-Vr).max()/np.abs(Vr).max():.2e} dense {np.abs(G@Wn-Vr).max()/np.abs(Vr).max():.2e}")
