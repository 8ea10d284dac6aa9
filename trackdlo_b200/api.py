"""ctypes binding of the C ABI (include/trackdlo_b200.h) used by tests/ and bench.py.

This is plumbing only: every call goes straight into libtrackdlo_b200.so (hand-written sm_100a
CUDA).  There is no CPU fallback -- if the library is missing or no B200 is present the calls
raise.  Argument names follow trackdlo::cpd_lle / trackdlo::tracking_step
(trackdlo/include/trackdlo.h:81-102).
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrackdlo_b200.so")

FE_EMPTY, FE_GRID, FE_CAPACITY = 1, 2, 4
ST_NOT_CONVERGED, ST_SINGULAR, ST_TOO_FEW_NODES, ST_EMPTY_CLOUD, ST_TRAVERSE_UB, ST_PRE_NOT_CONVERGED, ST_INVALID_INPUT = 1, 2, 4, 8, 16, 32, 64

ABI_SYMBOLS = [
    "tdlo_create", "tdlo_destroy", "tdlo_last_error", "tdlo_version",
    "tdlo_cpd_lle_batched", "tdlo_cpd_lle_batched_device",
    "tdlo_tracking_step_batched", "tdlo_tracking_step_batched_device",
    "tdlo_last_launch_info", "tdlo_synchronize", "tdlo_profile_phases", "tdlo_set_option",
    "tdlo_visibility_batched", "tdlo_visibility_batched_device", "tdlo_track_sequences", "tdlo_tracking_error_batched", "tdlo_tracking_error_batched_device",
    "tdlo_point_cloud_batched", "tdlo_point_cloud_batched_device", "tdlo_all_gather_packed",
]


class TdloError(RuntimeError):
    pass


class CpdParamsC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("beta", "lambda_", "lle_weight", "mu", "tol", "alpha", "k_vis",
                                           "visibility_threshold", "prune_radius")] + \
               [("max_iter", C.c_int32), ("include_lle", C.c_int32)]


class TrackParamsC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("visibility_threshold", "beta", "lambda_", "alpha", "k_vis", "mu", "tol",
                                           "beta_pre_proc", "lambda_pre_proc", "lle_weight", "prune_radius")] + \
               [("max_iter", C.c_int32), ("reserved", C.c_int32)]


class CpdBatchC(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("node_stride", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("X", "x_offsets", "n_nodes", "Y", "sigma2", "priors", "n_priors",
                                           "n_visible", "H", "W", "iters", "status")] + \
               [("priors_stride", C.c_int32), ("reserved", C.c_int32)]


class TrackBatchC(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_nodes", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("X", "x_offsets", "Y", "sigma2", "geodesic_coord", "visible",
                                           "visible_offsets", "visible_ext", "visible_ext_offsets", "H_pre",
                                           "guide_nodes", "priors", "n_priors", "iters", "status", "state", "packed_results")]


class VisBatchC(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_nodes", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("X", "x_offsets", "Y", "node_coord")] + \
               [("visibility_threshold", C.c_double), ("d_vis", C.c_double)] + \
               [(n, C.c_void_p) for n in ("dmin", "visible", "visible_offsets", "visible_ext", "visible_ext_offsets")] + \
               [("proj", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("pixel_width", C.c_int32), ("reserved", C.c_int32),
                ("not_self_occluded", C.c_void_p)]


class ErrBatchC(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_track", C.c_int32), ("n_true", C.c_int32), ("reserved", C.c_int32),
                ("Y_track", C.c_void_p), ("Y_true", C.c_void_p), ("error", C.c_void_p)]


class FrontendBatchC(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("multi_color", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("bgr", "depth", "occlusion_bgr", "proj")] + \
               [("hsv_lower", C.c_int32 * 3), ("hsv_upper", C.c_int32 * 3), ("leaf_size", C.c_double)] + \
               [("X", C.c_void_p), ("x_offsets", C.c_void_p), ("x_capacity", C.c_int64), ("status", C.c_void_p)]


class SeqBatchC(C.Structure):
    _fields_ = [("n_sequences", C.c_int32), ("n_nodes", C.c_int32), ("n_steps", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("X", "x_offsets", "Y", "sigma2", "geodesic_coord")] + [("d_vis", C.c_double)] + \
               [(n, C.c_void_p) for n in ("Y_traj", "iters_traj", "status_traj")] + \
               [("proj", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("pixel_width", C.c_int32), ("reserved", C.c_int32)]


@dataclass
class CpdParams:
    """Non-data arguments of trackdlo::cpd_lle (trackdlo.h:81-95); defaults = launch/trackdlo.launch."""
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
    prune_radius: float = 0.1

    def to_c(self):
        return CpdParamsC(self.beta, self.lambda_, self.lle_weight, self.mu, self.tol, self.alpha, self.k_vis,
                          self.visibility_threshold, self.prune_radius, self.max_iter, int(self.include_lle))


@dataclass
class TrackParams:
    """Constructor arguments of class trackdlo (trackdlo.h:59-71); defaults = launch/trackdlo.launch."""
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
    prune_radius: float = 0.1

    def to_c(self):
        return TrackParamsC(self.visibility_threshold, self.beta, self.lambda_, self.alpha, self.k_vis, self.mu,
                            self.tol, self.beta_pre_proc, self.lambda_pre_proc, self.lle_weight, self.prune_radius,
                            self.max_iter, 0)


_lib = None


def load_library():
    """Loads libtrackdlo_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TdloError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.tdlo_last_error.restype = C.c_char_p
        lib.tdlo_last_error.argtypes = [C.c_void_p]
        lib.tdlo_version.restype = C.c_char_p
        lib.tdlo_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int64]
        lib.tdlo_destroy.argtypes = [C.c_void_p]
        lib.tdlo_destroy.restype = None
        lib.tdlo_cpd_lle_batched.argtypes = [C.c_void_p, C.POINTER(CpdBatchC), C.POINTER(CpdParamsC)]
        lib.tdlo_cpd_lle_batched_device.argtypes = [C.c_void_p, C.POINTER(CpdBatchC), C.POINTER(CpdParamsC), C.c_void_p]
        lib.tdlo_tracking_step_batched.argtypes = [C.c_void_p, C.POINTER(TrackBatchC), C.POINTER(TrackParamsC)]
        lib.tdlo_tracking_step_batched_device.argtypes = [C.c_void_p, C.POINTER(TrackBatchC), C.POINTER(TrackParamsC), C.c_void_p]
        lib.tdlo_last_launch_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        lib.tdlo_synchronize.argtypes = [C.c_void_p]
        lib.tdlo_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        lib.tdlo_profile_phases.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
        lib.tdlo_tracking_error_batched.argtypes = [C.c_void_p, C.POINTER(ErrBatchC)]
        lib.tdlo_tracking_error_batched_device.argtypes = [C.c_void_p, C.POINTER(ErrBatchC), C.c_void_p]
        lib.tdlo_track_sequences.argtypes = [C.c_void_p, C.POINTER(SeqBatchC), C.POINTER(TrackParamsC)]
        lib.tdlo_point_cloud_batched.argtypes = [C.c_void_p, C.POINTER(FrontendBatchC)]
        lib.tdlo_point_cloud_batched_device.argtypes = [C.c_void_p, C.POINTER(FrontendBatchC), C.c_void_p]
        lib.tdlo_visibility_batched.argtypes = [C.c_void_p, C.POINTER(VisBatchC)]
        lib.tdlo_visibility_batched_device.argtypes = [C.c_void_p, C.POINTER(VisBatchC), C.c_void_p]
        _lib = lib
    return _lib


def _np(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a if shape is None else a.reshape(shape)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())          # torch tensor (host pinned or device)


class Context:
    """One tdlo_ctx (one GPU, one host thread)."""

    def __init__(self, max_frames, max_nodes, max_points_total, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.tdlo_create(C.byref(self.h), device, max_frames, max_nodes, max_points_total)
        if rc != 0:
            raise TdloError(f"tdlo_create failed ({rc}): {self.lib.tdlo_last_error(None).decode()}")
        self.max_frames, self.max_nodes = max_frames, max_nodes

    def close(self):
        if self.h:
            self.lib.tdlo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise TdloError(f"{what} failed ({rc}): {self.lib.tdlo_last_error(self.h).decode()}")

    def synchronize(self):
        """tdlo_synchronize: waits for the last *_device call; raises if the kernel's watchdog gave up."""
        self._check(self.lib.tdlo_synchronize(self.h), "tdlo_synchronize")

    OPTIONS = {"chunk_points": 2, "truncation": 3, "inflight": 4, "threads": 5, "watchdog_ms": 6, "solver": 7, "voxel_cells": 8, "truncation_rel": 9}

    def set_option(self, name, value):
        """tdlo_set_option: chunk_points, truncation, inflight, threads, watchdog_ms."""
        self._check(self.lib.tdlo_set_option(self.h, self.OPTIONS[name], float(value)), f"tdlo_set_option({name})")

    def profile_phases(self, enable=True):
        """Returns and resets the kernel's phase cycle counters; see tdlo_profile_phases."""
        cyc = (C.c_uint64 * 16)()
        self._check(self.lib.tdlo_profile_phases(self.h, int(enable), cyc), "tdlo_profile_phases")
        v = list(cyc)
        return {"cycles": dict(zip(self.PHASE_CYCLES, v[:10])), "counts": dict(zip(self.PHASE_COUNTS, v[10:16]))}

    # slot meaning of tdlo_profile_phases (include/trackdlo_b200.h): thread-0 cycles summed over all CTAs, then counters
    PHASE_CYCLES = ("queue_wait", "prune", "visibility_prepass", "estep", "start_call", "wave_glue", "mstep_gather_assemble",
                    "solve", "update", "finish_call")
    PHASE_COUNTS = ("estep_tasks", "tiles", "window_rows", "row_blocks", "warp_tile_loop_cycles", "end_of_task_reduction_cycles")

    def launch_info(self):
        info = (C.c_int32 * 8)()
        self.lib.tdlo_last_launch_info(self.h, info)
        keys = ("reserved", "ctas", "threads", "smem_bytes", "tile_points", "launches", "ctas_per_sm", "sm_count")
        return dict(zip(keys, list(info)))

    # ------------------------------------------------------------------ cpd_lle, host buffers
    def cpd_lle_batched(self, X, x_offsets, Y, sigma2, params: CpdParams, n_nodes=None, priors=None, n_priors=None,
                        n_visible=None, H=None):
        """X [sum Mp,3], x_offsets [F+1], Y [F,S,3], sigma2 [F], priors [F,PS,4] (PS = any stride >= max n_priors)
        ->  dict(Y, sigma2, W, iters, status)."""
        X = _np(X, np.float64, (-1, 3)); xo = _np(x_offsets, np.int64)
        Y = _np(Y, np.float64).copy(); F, S = Y.shape[0], Y.shape[1]
        s2 = _np(sigma2, np.float64).copy().reshape(F)
        W = np.zeros((F, S, 3)); iters = np.zeros(F, np.int32); status = np.zeros(F, np.int32)
        nn = None if n_nodes is None else _np(n_nodes, np.int32)
        pr = None if priors is None else _np(priors, np.float64, (F, -1, 4))
        npr = None if n_priors is None else _np(n_priors, np.int32)
        nv = None if n_visible is None else _np(n_visible, np.int32)
        Hc = None if H is None else _np(H, np.float64, (F, S, S))
        b = CpdBatchC(F, S, _ptr(X), _ptr(xo), _ptr(nn), _ptr(Y), _ptr(s2), _ptr(pr), _ptr(npr), _ptr(nv), _ptr(Hc),
                      _ptr(W), _ptr(iters), _ptr(status), 0 if pr is None else pr.shape[1], 0)
        pc = params.to_c()
        self._check(self.lib.tdlo_cpd_lle_batched(self.h, C.byref(b), C.byref(pc)), "tdlo_cpd_lle_batched")
        return dict(Y=Y, sigma2=s2, W=W, iters=iters, status=status)

    def cpd_lle_batched_raw(self, batch: CpdBatchC, params: CpdParamsC, device=False, stream=0):
        """Thin call with caller-owned buffers (pinned host or device pointers); used by bench.py."""
        if device:
            rc = self.lib.tdlo_cpd_lle_batched_device(self.h, C.byref(batch), C.byref(params), C.c_void_p(stream))
        else:
            rc = self.lib.tdlo_cpd_lle_batched(self.h, C.byref(batch), C.byref(params))
        self._check(rc, "tdlo_cpd_lle_batched")

    # ------------------------------------------------------------------ tracking_step, host buffers
    def tracking_step_batched(self, X, x_offsets, Y, sigma2, geodesic_coord, visible, visible_offsets, visible_ext,
                              visible_ext_offsets, params: TrackParams, H_pre=None):
        X = _np(X, np.float64, (-1, 3)); xo = _np(x_offsets, np.int64)
        Y = _np(Y, np.float64).copy(); F, N = Y.shape[0], Y.shape[1]
        s2 = _np(sigma2, np.float64).copy().reshape(F)
        geo = _np(geodesic_coord, np.float64, (F, N))
        vis = _np(visible, np.int32); vo = _np(visible_offsets, np.int64)
        ext = _np(visible_ext, np.int32); eo = _np(visible_ext_offsets, np.int64)
        Hc = None if H_pre is None else _np(H_pre, np.float64, (F, N, N))
        guide = np.zeros((F, N, 3)); pri = np.zeros((F, 2 * N, 4)); npri = np.zeros(F, np.int32)
        iters = np.zeros((F, 2), np.int32); status = np.zeros(F, np.int32); state = np.zeros(F, np.int32)
        packed = np.zeros((F, 3 * N + 4))
        b = TrackBatchC(F, N, _ptr(X), _ptr(xo), _ptr(Y), _ptr(s2), _ptr(geo), _ptr(vis), _ptr(vo), _ptr(ext), _ptr(eo),
                        _ptr(Hc), _ptr(guide), _ptr(pri), _ptr(npri), _ptr(iters), _ptr(status), _ptr(state), _ptr(packed))
        pc = params.to_c()
        self._check(self.lib.tdlo_tracking_step_batched(self.h, C.byref(b), C.byref(pc)), "tdlo_tracking_step_batched")
        return dict(Y=Y, sigma2=s2, guide=guide, priors=pri, n_priors=npri, iters=iters, status=status, state=state, packed=packed)

    # ------------------------------------------------------------------ visibility front-end (trackdlo_node.cpp:254-277, 346-360)
    def visibility_batched(self, X, x_offsets, Y, node_coord, visibility_threshold=0.008, d_vis=0.06, proj=None, rows=0, cols=0,
                           pixel_width=40):
        """proj [F,3,4] (with rows, cols, pixel_width) switches the self-occlusion test of trackdlo_node.cpp:280-343 on."""
        X = _np(X, np.float64, (-1, 3)); xo = _np(x_offsets, np.int64)
        Y = _np(Y, np.float64); F, N = Y.shape[0], Y.shape[1]
        nc = _np(node_coord, np.float64, (F, N))
        dmin = np.zeros((F, N)); vis = np.zeros(F * N, np.int32); ext = np.zeros(F * N, np.int32)
        vo = np.zeros(F + 1, np.int64); eo = np.zeros(F + 1, np.int64)
        pr = None if proj is None else _np(proj, np.float64, (F, 12))
        free = np.ones((F, N), np.int32)
        b = VisBatchC(F, N, _ptr(X), _ptr(xo), _ptr(Y), _ptr(nc), visibility_threshold, d_vis, _ptr(dmin), _ptr(vis), _ptr(vo), _ptr(ext), _ptr(eo),
                      _ptr(pr), int(rows), int(cols), int(pixel_width), 0, _ptr(free) if pr is not None else None)
        self._check(self.lib.tdlo_visibility_batched(self.h, C.byref(b)), "tdlo_visibility_batched")
        return dict(dmin=dmin, visible=vis[:vo[F]].copy(), visible_offsets=vo, visible_ext=ext[:eo[F]].copy(), visible_ext_offsets=eo,
                    not_self_occluded=free)

    # ------------------------------------------------------------------ perception front-end (trackdlo_node.cpp:159-242)
    def point_cloud_batched(self, bgr, depth, proj, hsv_lower=(90, 90, 30), hsv_upper=(130, 255, 255), multi_color=False,
                            occlusion_bgr=None, leaf_size=0.008, x_capacity=None):
        """bgr [F,H,W,3] uint8, depth [F,H,W] uint16 (mm), proj [F,3,4]  ->  dict(X [sum Mp,3], x_offsets [F+1], status [F])."""
        bgr = _np(bgr, np.uint8); F, H, W = bgr.shape[0], bgr.shape[1], bgr.shape[2]
        depth = _np(depth, np.uint16, (F, H, W)); proj = _np(proj, np.float64, (F, 12))
        occ = None if occlusion_bgr is None else _np(occlusion_bgr, np.uint8, (F, H, W, 3))
        cap = int(x_capacity if x_capacity is not None else F * H * W)
        X = np.zeros((cap, 3)); xo = np.zeros(F + 1, np.int64); st = np.zeros(F, np.int32)
        b = FrontendBatchC(F, H, W, int(multi_color), _ptr(bgr), _ptr(depth), _ptr(occ), _ptr(proj), (C.c_int32 * 3)(*hsv_lower),
                           (C.c_int32 * 3)(*hsv_upper), leaf_size, _ptr(X), _ptr(xo), cap, _ptr(st))
        self._check(self.lib.tdlo_point_cloud_batched(self.h, C.byref(b)), "tdlo_point_cloud_batched")
        return dict(X=X[:xo[F]].copy(), x_offsets=xo, status=st)

    def point_cloud_batched_raw(self, batch: "FrontendBatchC", device=True, stream=0):
        if device:
            rc = self.lib.tdlo_point_cloud_batched_device(self.h, C.byref(batch), C.c_void_p(stream))
        else:
            rc = self.lib.tdlo_point_cloud_batched(self.h, C.byref(batch))
        self._check(rc, "tdlo_point_cloud_batched")

    # ------------------------------------------------------------------ evaluator frame error (evaluator.cpp:233-283, 333-341)
    def tracking_error_batched(self, Y_track, Y_true):
        a = _np(Y_track, np.float64); b = _np(Y_true, np.float64)
        F = a.shape[0]
        err = np.zeros(F)
        eb = ErrBatchC(F, a.shape[1], b.shape[1], 0, _ptr(a), _ptr(b), _ptr(err))
        self._check(self.lib.tdlo_tracking_error_batched(self.h, C.byref(eb)), "tdlo_tracking_error_batched")
        return err

    # ------------------------------------------------------------------ sequence mode (visibility + tracking_step per frame on the device)
    def track_sequences(self, X, x_offsets, Y, sigma2, geodesic_coord, params: TrackParams, n_steps, d_vis=0.06, proj=None, rows=0, cols=0,
                        pixel_width=40):
        """proj [S,3,4] (with rows, cols, pixel_width) adds the self-occlusion test to every step's visibility lists."""
        X = _np(X, np.float64, (-1, 3)); xo = _np(x_offsets, np.int64)
        Y = _np(Y, np.float64).copy(); S, N = Y.shape[0], Y.shape[1]
        s2 = _np(sigma2, np.float64).copy().reshape(S)
        geo = _np(geodesic_coord, np.float64, (S, N))
        traj = np.zeros((n_steps, S, N, 3)); its = np.zeros((n_steps, S, 2), np.int32); st = np.zeros((n_steps, S), np.int32)
        pr = None if proj is None else _np(proj, np.float64, (S, 12))
        b = SeqBatchC(S, N, n_steps, _ptr(X), _ptr(xo), _ptr(Y), _ptr(s2), _ptr(geo), d_vis, _ptr(traj), _ptr(its), _ptr(st),
                      _ptr(pr), int(rows), int(cols), int(pixel_width), 0)
        pc = params.to_c()
        self._check(self.lib.tdlo_track_sequences(self.h, C.byref(b), C.byref(pc)), "tdlo_track_sequences")
        return dict(Y=Y, sigma2=s2, Y_traj=traj, iters=its, status=st)

    def visibility_batched_raw(self, batch: VisBatchC, device=True, stream=0):
        if device:
            rc = self.lib.tdlo_visibility_batched_device(self.h, C.byref(batch), C.c_void_p(stream))
        else:
            rc = self.lib.tdlo_visibility_batched(self.h, C.byref(batch))
        self._check(rc, "tdlo_visibility_batched")

    def tracking_step_batched_raw(self, batch: TrackBatchC, params: TrackParamsC, device=False, stream=0):
        if device:
            rc = self.lib.tdlo_tracking_step_batched_device(self.h, C.byref(batch), C.byref(params), C.c_void_p(stream))
        else:
            rc = self.lib.tdlo_tracking_step_batched(self.h, C.byref(batch), C.byref(params))
        self._check(rc, "tdlo_tracking_step_batched")
