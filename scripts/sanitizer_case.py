"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): gpurun -- compute-sanitizer --tool racecheck python scripts/sanitizer_case.py"""
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from trackdlo_b200 import api, synth
f = synth.make_frame(0, n_nodes=30, n_points=700, occlusion=0.3)
ctx = api.Context(max_frames=2, max_nodes=30, max_points_total=1400)
one = lambda n: np.array([0, n, 2 * n], np.int64)
X = np.concatenate([f["X"], f["X"]]); n = len(f["X"])
Y = np.stack([f["Y"], f["Y"]]); rest = np.stack([f["rest"], f["rest"]])
vis = np.concatenate([f["vis"], f["vis"]]); ext = np.concatenate([f["vis_ext"], f["vis_ext"]])
vo = np.array([0, len(f["vis"]), 2 * len(f["vis"])], np.int64); eo = np.array([0, len(f["vis_ext"]), 2 * len(f["vis_ext"])], np.int64)
r = ctx.tracking_step_batched(X, one(n), Y, np.zeros(2), rest, vis, vo, ext, eo, api.TrackParams(max_iter=4))
print("iters", r["iters"].tolist(), "status", r["status"].tolist())
v = ctx.visibility_batched(X, one(n), Y, rest)
print("vis ok", len(v["visible"]))
ctx.close()
# Nn = 100: blocked Cholesky (16-column panels, FP64 MMA); Nn = 200: 8/16-column panels
for Nn, Mp in ((100, 1500), (200, 2500)):
    g = synth.make_frame(1, n_nodes=Nn, n_points=Mp)
    c2 = api.Context(max_frames=1, max_nodes=Nn, max_points_total=Mp)
    r2 = c2.cpd_lle_batched(g["X"], np.array([0, Mp], np.int64), g["Y"][None], np.zeros(1), api.CpdParams(max_iter=2, tol=0.0))
    print("Nn", Nn, "iters", r2["iters"].tolist(), "status", r2["status"].tolist())
    c2.close()
# tracking states 2 (tail occluded) and 4 (both ends occluded -> traverse_euclidean alignment 2)
for win in ([(0.7, 1.0)], [(0.0, 0.2), (0.8, 1.0)]):
    g = synth.make_frame(1, n_nodes=40, n_points=1500, occl_windows=win)
    c3 = api.Context(max_frames=1, max_nodes=40, max_points_total=len(g["X"]))
    o1 = lambda m: np.array([0, m], np.int64)
    r3 = c3.tracking_step_batched(g["X"], o1(len(g["X"])), g["Y"][None], np.zeros(1), g["rest"][None], g["vis"], o1(len(g["vis"])), g["vis_ext"], o1(len(g["vis_ext"])), api.TrackParams(max_iter=4))
    print("state", r3["state"].tolist(), "iters", r3["iters"].tolist(), "status", r3["status"].tolist())
    c3.close()
# structured solves: Nn = 100 tracking_step (LLE state-space solve above 64 nodes), and solver = 2 at Nn = 30
g = synth.make_frame(2, n_nodes=100, n_points=1800)
c4 = api.Context(max_frames=1, max_nodes=100, max_points_total=len(g["X"]))
o1 = lambda m: np.array([0, m], np.int64)
r4 = c4.tracking_step_batched(g["X"], o1(len(g["X"])), g["Y"][None], np.zeros(1), g["rest"][None], g["vis"], o1(len(g["vis"])), g["vis_ext"], o1(len(g["vis_ext"])), api.TrackParams(max_iter=3))
print("Nn=100 tracking iters", r4["iters"].tolist(), "status", r4["status"].tolist())
c4.close()
c5 = api.Context(max_frames=1, max_nodes=30, max_points_total=700)
c5.set_option("solver", 2)
r5 = c5.tracking_step_batched(f["X"], o1(n), f["Y"][None], np.zeros(1), f["rest"][None], f["vis"], o1(len(f["vis"])), f["vis_ext"], o1(len(f["vis_ext"])), api.TrackParams(max_iter=3))
print("solver=2 iters", r5["iters"].tolist())
# single-chunk inline path (<= 256 points) and the perception front-end
h = synth.make_frame(3, n_nodes=30, n_points=200)
r6 = c5.tracking_step_batched(h["X"], o1(200), h["Y"][None], np.zeros(1), h["rest"][None], h["vis"], o1(len(h["vis"])), h["vis_ext"], o1(len(h["vis_ext"])), api.TrackParams(max_iter=3))
print("inline iters", r6["iters"].tolist())
fr = synth.render_frame(0, rows=90, cols=160, width_px=3)
pc = c5.point_cloud_batched(fr["bgr"][None], fr["depth"][None], fr["proj"][None])
print("front-end points", int(pc["x_offsets"][1]), "status", pc["status"].tolist())
c5.close()
# visibility lists with the self-occlusion test (thick-line coverage of projected edges)
import glob, os
gs = np.load(sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "selfocc_coil.npz")))[0])
c6 = api.Context(max_frames=1, max_nodes=len(gs["Y"]), max_points_total=len(gs["X"]))
rv = c6.visibility_batched(gs["X"], o1(len(gs["X"])), gs["Y"][None], gs["node_coord"][None], proj=gs["proj"][None], rows=int(gs["rows"]), cols=int(gs["cols"]),
                           pixel_width=int(gs["pixel_width"]))
print("self-occluded nodes", int((rv["not_self_occluded"] == 0).sum()), "visible", len(rv["visible"]))
c6.close()
