"""ctypes loader for oracle/_ref/libtrackdlo_ref.so: the UNMODIFIED reference sources
(/root/reference/trackdlo/src/trackdlo.cpp + utils.cpp) compiled by oracle/Makefile (target `_ref`) behind
oracle/ref_harness.cpp.  Test infrastructure: it pins oracle/trackdlo_oracle.cpp (tests/test_ref_pin.py) and checks
the CUDA path directly on the GPU box (the prebuilt .so travels there; /root/reference does not).

Same call shapes as oracle/oracle.py so a test can swap one for the other."""
import ctypes as C
import os

import numpy as np

from .oracle import CpdParams, TrackParams, _CpdP, _TrackP, _f64, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libtrackdlo_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libtrackdlo_ref.so is missing: run `make -C oracle _ref` where /root/reference exists")
        _lib = C.CDLL(_SO)
        _lib.ref_cpd_lle.restype = C.c_int
        _lib.ref_tracking_step.restype = C.c_int
        _lib.ref_traverse_euclidean.restype = C.c_int
        _lib.ref_line_sphere_intersection.restype = C.c_int
        _lib.ref_pt2pt_dis.restype = C.c_double
        _lib.ref_pt2pt_dis_sq.restype = C.c_double
        _lib.ref_uses_eigen_shim.restype = C.c_int
        _lib.ref_lle.restype = None
        _lib.ref_eigen_cod_solve.restype = None
        _lib.ref_eigen_inverse.restype = C.c_double
    return _lib


def uses_eigen_shim():
    return bool(lib().ref_uses_eigen_shim())


def cpd_lle(X, Y, sigma2, prm: CpdParams, priors=None, vis=None):
    """trackdlo::cpd_lle of the reference.  Returns dict(Y, sigma2, iters, converged)."""
    L = lib()
    X = _f64(X, (-1, 3)); Y = _f64(Y, (-1, 3)).copy()
    pr = _f64(priors if priors is not None else np.zeros((0, 4)), (-1, 4))
    vs = np.ascontiguousarray(vis if vis is not None else np.zeros(0), dtype=np.int32)
    cp = _CpdP(prm.beta, prm.lambda_, prm.lle_weight, prm.mu, prm.tol, prm.alpha, prm.k_vis,
               prm.visibility_threshold, prm.max_iter, int(prm.include_lle))
    s2 = C.c_double(sigma2); it = C.c_int32(0)
    conv = L.ref_cpd_lle(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int32(Y.shape[0]), C.byref(s2), C.byref(cp),
                         _p(pr), C.c_int32(pr.shape[0]), _p(vs), C.c_int32(vs.shape[0]), C.byref(it))
    return dict(Y=Y, sigma2=s2.value, iters=it.value, converged=bool(conv))


def tracking_step(X, Y, sigma2, geodesic_coord, vis, vis_ext, tp: TrackParams):
    """The reference's tracking_step driven as trackdlo_node.cpp drives it.
    Returns dict(Y, sigma2, guide, priors, iters, state)."""
    L = lib()
    X = _f64(X, (-1, 3)); Y = _f64(Y, (-1, 3)).copy(); Nn = Y.shape[0]
    geo = _f64(geodesic_coord)
    v = np.ascontiguousarray(vis, dtype=np.int32); ve = np.ascontiguousarray(vis_ext, dtype=np.int32)
    t = _TrackP(tp.visibility_threshold, tp.beta, tp.lambda_, tp.alpha, tp.k_vis, tp.mu, tp.tol,
                tp.beta_pre_proc, tp.lambda_pre_proc, tp.lle_weight, tp.max_iter, 0)
    s2 = C.c_double(sigma2)
    guide = np.zeros((len(ve), 3)); pri = np.zeros((2 * Nn + 2, 4)); npri = C.c_int32(0)
    its = np.zeros(2, np.int32); st = C.c_int32(-1)
    L.ref_tracking_step(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int32(Nn), C.byref(s2), _p(geo),
                        _p(v), C.c_int32(len(v)), _p(ve), C.c_int32(len(ve)), C.byref(t),
                        _p(guide), _p(pri), C.byref(npri), _p(its), C.byref(st))
    return dict(Y=Y, sigma2=s2.value, guide=guide, priors=pri[:npri.value].copy(), iters=its, state=st.value)


def traverse_euclidean(geodesic_coord, guide, vis, alignment, align_idx=-1):
    L = lib()
    geo = _f64(geodesic_coord); g = _f64(guide, (-1, 3)); v = np.ascontiguousarray(vis, dtype=np.int32)
    out = np.zeros((len(geo) + 2, 4)); n = C.c_int32(0)
    L.ref_traverse_euclidean(_p(geo), C.c_int32(len(geo)), _p(g), C.c_int32(g.shape[0]), _p(v), C.c_int32(len(v)),
                             C.c_int32(alignment), C.c_int32(align_idx), _p(out), C.byref(n))
    return out[:n.value].copy()


def lle(Y):
    """(L, H) of trackdlo::calc_LLE_weights(6, Y) and trackdlo.cpp:237."""
    Y = _f64(Y, (-1, 3)); Nn = Y.shape[0]
    Lm = np.zeros((Nn, Nn)); H = np.zeros((Nn, Nn))
    lib().ref_lle(_p(Y), C.c_int32(Nn), _p(Lm), _p(H))
    return Lm, H


def line_sphere_intersection(A, B, centre, radius):
    out = np.zeros((2, 3))
    n = lib().ref_line_sphere_intersection(_p(_f64(A)), _p(_f64(B)), _p(_f64(centre)), C.c_double(radius), _p(out))
    return out[:n].copy()


def pt2pt_dis(a, b):
    a = _f64(a, (-1, 3)); b = _f64(b, (-1, 3))
    return float(lib().ref_pt2pt_dis(_p(a), _p(b), C.c_int32(a.shape[0])))


def pt2pt_dis_sq(a, b):
    a = _f64(a, (-1, 3)); b = _f64(b, (-1, 3))
    return float(lib().ref_pt2pt_dis_sq(_p(a), _p(b), C.c_int32(a.shape[0])))


def eigen_cod_solve(A, B):
    """A.completeOrthogonalDecomposition().solve(B) as the reference build computes it (trackdlo.cpp:415)."""
    A = _f64(A); n = A.shape[0]; B = _f64(B, (n, -1)); X = np.zeros((A.shape[1], B.shape[1]))
    assert A.shape[0] == A.shape[1]
    lib().ref_eigen_cod_solve(_p(A), _p(B), C.c_int32(n), C.c_int32(B.shape[1]), _p(X))
    return X


def eigen_inverse(A):
    """(A.inverse(), A.determinant()) as the reference build computes them (trackdlo.cpp:136-143)."""
    A = _f64(A); n = A.shape[0]; inv = np.zeros((n, n))
    det = lib().ref_eigen_inverse(_p(A), C.c_int32(n), _p(inv))
    return inv, float(det)
