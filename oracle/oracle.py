"""ctypes loader for oracle/liboracle.so (the C++17 restatement of trackdlo.cpp:161-441, 584-999)."""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "trackdlo_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


class _CpdP(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("beta", "lambda_", "lle_weight", "mu", "tol", "alpha", "k_vis",
                                           "visibility_threshold")] + [("max_iter", C.c_int32), ("include_lle", C.c_int32)]


class _Trace(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("P1", "PX", "Np", "sigma2", "W", "Y", "A", "B")]


class _TrackP(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("visibility_threshold", "beta", "lambda_", "alpha", "k_vis", "mu", "tol",
                                           "beta_pre_proc", "lambda_pre_proc", "lle_weight")] + \
               [("max_iter", C.c_int32), ("pad_", C.c_int32)]


@dataclass
class CpdParams:
    """Arguments of trackdlo::cpd_lle (trackdlo.h:81-95)."""
    beta: float = 0.35
    lambda_: float = 50000.0
    lle_weight: float = 10.0
    mu: float = 0.1
    max_iter: int = 50
    tol: float = 0.0002
    include_lle: bool = False
    alpha: float = 0.0
    k_vis: float = 0.0
    visibility_threshold: float = 0.01


@dataclass
class TrackParams:
    """Constructor arguments of trackdlo (trackdlo.h:59-71); defaults = launch/trackdlo.launch."""
    visibility_threshold: float = 0.008
    beta: float = 0.35
    lambda_: float = 50000.0
    alpha: float = 3.0
    k_vis: float = 50.0
    mu: float = 0.1
    max_iter: int = 50
    tol: float = 0.0002
    beta_pre_proc: float = 3.0
    lambda_pre_proc: float = 1.0
    lle_weight: float = 10.0


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_cpd_lle.restype = C.c_int
        _lib.oracle_tracking_step.restype = C.c_int
        _lib.oracle_traverse_euclidean.restype = C.c_int
        _lib.oracle_lle_H.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def cpd_lle(X, Y, sigma2, prm: CpdParams, priors=None, vis=None, H=None, trace=False):
    """Runs the oracle's cpd_lle.  Returns dict(Y, sigma2, W, iters, converged, kept[, trace])."""
    L = lib()
    X = _f64(X, (-1, 3)); Y = _f64(Y, (-1, 3)).copy()
    Nn = Y.shape[0]
    pr = _f64(priors if priors is not None else np.zeros((0, 4)), (-1, 4))
    vs = np.ascontiguousarray(vis if vis is not None else np.zeros(0), dtype=np.int32)
    Hc = None if H is None else _f64(H, (Nn, Nn))
    cp = _CpdP(prm.beta, prm.lambda_, prm.lle_weight, prm.mu, prm.tol, prm.alpha, prm.k_vis,
               prm.visibility_threshold, prm.max_iter, int(prm.include_lle))
    s2 = C.c_double(sigma2)
    W = np.zeros((Nn, 3)); it = C.c_int32(0); kept = C.c_int64(0)
    tr = None; trs = None
    if trace:
        K = max(prm.max_iter, 1)
        tr = dict(P1=np.zeros((K, Nn)), PX=np.zeros((K, Nn, 3)), Np=np.zeros(K), sigma2=np.zeros(K),
                  W=np.zeros((K, Nn, 3)), Y=np.zeros((K, Nn, 3)), A=np.zeros((Nn, Nn)), B=np.zeros((Nn, 3)))
        trs = _Trace(*[_p(tr[k]) for k in ("P1", "PX", "Np", "sigma2", "W", "Y", "A", "B")])
    conv = L.oracle_cpd_lle(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int32(Nn), C.byref(s2), C.byref(cp),
                            _p(pr), C.c_int32(pr.shape[0]), _p(vs), C.c_int32(vs.shape[0]), _p(Hc),
                            _p(W), C.byref(it), C.byref(kept), C.byref(trs) if trs is not None else None)
    out = dict(Y=Y, sigma2=s2.value, W=W, iters=it.value, converged=bool(conv), kept=kept.value)
    if trace:
        out["trace"] = {k: (v[:it.value] if k not in ("A", "B") else v) for k, v in tr.items()}
    return out


def lle_H(Y):
    Y = _f64(Y, (-1, 3)); Nn = Y.shape[0]
    H = np.zeros((Nn, Nn))
    lib().oracle_lle_H(_p(Y), C.c_int32(Nn), _p(H))
    return H


def tracking_step(X, Y, sigma2, geodesic_coord, vis, vis_ext, tp: TrackParams, H=None):
    """Oracle tracking_step (trackdlo.cpp:900-999).  Returns dict(Y, sigma2, guide, priors, iters, converged, state, err)."""
    L = lib()
    X = _f64(X, (-1, 3)); Y = _f64(Y, (-1, 3)).copy(); Nn = Y.shape[0]
    geo = _f64(geodesic_coord)
    v = np.ascontiguousarray(vis, dtype=np.int32); ve = np.ascontiguousarray(vis_ext, dtype=np.int32)
    Hc = None if H is None else _f64(H, (len(ve), len(ve)))
    t = _TrackP(tp.visibility_threshold, tp.beta, tp.lambda_, tp.alpha, tp.k_vis, tp.mu, tp.tol,
                tp.beta_pre_proc, tp.lambda_pre_proc, tp.lle_weight, tp.max_iter, 0)
    s2 = C.c_double(sigma2)
    guide = np.zeros((len(ve), 3)); pri = np.zeros((2 * Nn + 2, 4)); npri = C.c_int32(0)
    its = np.zeros(2, np.int32); cv = np.zeros(2, np.int32); st = C.c_int32(-1)
    err = L.oracle_tracking_step(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int32(Nn), C.byref(s2), _p(geo),
                                 _p(v), C.c_int32(len(v)), _p(ve), C.c_int32(len(ve)), C.byref(t), _p(Hc),
                                 _p(guide), _p(pri), C.byref(npri), _p(its), _p(cv), C.byref(st))
    return dict(Y=Y, sigma2=s2.value, guide=guide, priors=pri[:npri.value].copy(), iters=its, converged=cv,
                state=st.value, err=err)


def visibility(X, Y, node_coord, visibility_threshold=0.008, d_vis=0.06):
    """Oracle of the visibility front-end (trackdlo_node.cpp:254-277, 346-360, raster excluded).
    Returns dict(dmin [Nn], vis, vis_ext)."""
    L = lib()
    X = _f64(X, (-1, 3)); Y = _f64(Y, (-1, 3)); Nn = Y.shape[0]
    nc = _f64(node_coord)
    dmin = np.zeros(Nn); vis = np.zeros(Nn, np.int32); ext = np.zeros(Nn, np.int32); ne = C.c_int32(0)
    L.oracle_visibility.restype = C.c_int
    nv = L.oracle_visibility(_p(X), C.c_int64(X.shape[0]), _p(Y), C.c_int32(Nn), _p(nc), C.c_double(visibility_threshold),
                             C.c_double(d_vis), _p(dmin), _p(vis), _p(ext), C.byref(ne))
    return dict(dmin=dmin, vis=vis[:nv].copy(), vis_ext=ext[:ne.value].copy())


def tracking_error(Y_track, Y_true):
    """Oracle of the evaluator's frame error (evaluator.cpp:233-283, 333-341)."""
    L = lib()
    a = _f64(Y_track, (-1, 3)); b = _f64(Y_true, (-1, 3))
    L.oracle_tracking_error.restype = C.c_double
    return float(L.oracle_tracking_error(_p(a), C.c_int32(a.shape[0]), _p(b), C.c_int32(b.shape[0])))


def traverse_euclidean(geodesic_coord, guide, vis, alignment, align_idx=-1):
    L = lib()
    geo = _f64(geodesic_coord); g = _f64(guide, (-1, 3)); v = np.ascontiguousarray(vis, dtype=np.int32)
    out = np.zeros((len(geo) + 2, 4)); n = C.c_int32(0)
    err = L.oracle_traverse_euclidean(_p(geo), C.c_int32(len(geo)), _p(g), C.c_int32(g.shape[0]), _p(v),
                                      C.c_int32(len(v)), C.c_int32(alignment), C.c_int32(align_idx), _p(out), C.byref(n))
    return out[:n.value].copy(), err
