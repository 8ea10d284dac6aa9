"""The reference's live regime (one DLO, ~45 nodes, a few hundred points after the 8 mm voxel grid, default
tol): per-frame latency of tracking_step (a) through the C ABI with host buffers, batch of one, (b) through the drop-in
`class trackdlo` adapter (MatrixXd in/out, tests/cpp/adapter_latency.cpp), (c) the oracle on one pinned host core, and
(d) the reference's own sources (oracle/_ref, Eigen stand-in) -- same frame, same parameters.  One JSON line per size.
Usage (GPU box): python scripts/production_regime.py > profiles/r2_production_regime.jsonl"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from oracle import ref
from trackdlo_b200 import api, synth

N = 45
tp, otp = api.TrackParams(), oracle.TrackParams()
one = lambda n: np.array([0, n], np.int64)
with tempfile.TemporaryDirectory() as d:
    exe = os.path.join(d, "adapter_latency")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"),
                           os.path.join(ROOT, "tests", "cpp", "adapter_latency.cpp"), "-o", exe, "-L", os.path.join(ROOT, "trackdlo_b200"),
                           "-ltrackdlo_b200", "-Wl,-rpath," + os.path.join(ROOT, "trackdlo_b200")])
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    for Mp in (300, 1000, 3000):
        f = synth.make_frame(5, n_nodes=N, n_points=Mp, tau_vis=0.02 if Mp < 1000 else 0.008)
        ctx = api.Context(max_frames=1, max_nodes=64, max_points_total=max(Mp, 256))
        args = (f["X"], one(len(f["X"])), f["Y"][None], np.zeros(1), f["rest"][None], f["vis"], one(len(f["vis"])), f["vis_ext"], one(len(f["vis_ext"])), tp)
        for _ in range(5):
            r = ctx.tracking_step_batched(*args)
        t0 = time.perf_counter(); K = 200
        for _ in range(K):
            r = ctx.tracking_step_batched(*args)
        abi_ms = (time.perf_counter() - t0) / K * 1e3
        info = ctx.launch_info()
        ctx.close()
        t0 = time.perf_counter(); Ko = 20
        for _ in range(Ko):
            o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], otp)
        cpu_ms = (time.perf_counter() - t0) / Ko * 1e3
        ref_ms = None
        if ref.available():
            t0 = time.perf_counter()
            for _ in range(5):
                ref.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], otp)
            ref_ms = (time.perf_counter() - t0) / 5 * 1e3
        fin = os.path.join(d, "in.bin")
        with open(fin, "wb") as fh:
            np.array([N, len(f["X"]), len(f["vis"]), len(f["vis_ext"])], np.int64).tofile(fh)
            np.array([tp.visibility_threshold, tp.beta, tp.lambda_, tp.alpha, tp.k_vis, tp.mu, tp.max_iter, tp.tol, tp.beta_pre_proc, tp.lambda_pre_proc, tp.lle_weight, 0.0]).tofile(fh)
            f["Y"].tofile(fh); f["rest"].tofile(fh); f["X"].tofile(fh); f["vis"].astype(np.int32).tofile(fh); f["vis_ext"].astype(np.int32).tofile(fh)
        ad = json.loads(subprocess.check_output([exe, fin, "200"]).decode().strip().split("\n")[-1])
        rel = float(np.abs(r["Y"][0] - o["Y"]).max() / np.abs(o["Y"]).max())
        print(json.dumps({"nodes": N, "points": Mp, "iters": [int(v) for v in r["iters"][0]], "iters_oracle": [int(v) for v in o["iters"]], "rel_err_Y": rel,
                          "gpu_c_abi_ms": abi_ms, **ad, "cpu_oracle_1core_ms": cpu_ms, "cpu_reference_build_1core_ms": ref_ms,
                          "chunk_points": info["tile_points"], "note": "c_abi timing includes the ctypes/NumPy marshalling of api.py (~0.05 ms)"}), flush=True)
