"""Development aid: times the C2 workload (device-resident) for several cluster sizes and prints the
kernel's per-phase cycle counters."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from trackdlo_b200 import api, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "track"
clusters = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2, 4]
NODES = int(sys.argv[4]) if len(sys.argv) > 4 else 50
POINTS = int(sys.argv[5]) if len(sys.argv) > 5 else 20000
wl = synth.make_batch(F, n_nodes=NODES, n_points=POINTS)
dev = torch.device("cuda:0")
ctx = api.Context(max_frames=F, max_nodes=NODES, max_points_total=int(wl["x_offsets"][-1]))
d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y", "rest", "vis", "vis_offsets", "vis_ext", "vis_ext_offsets")}
Y0 = d["Y"].clone()
s2 = torch.zeros(F, dtype=torch.float64, device=dev)
it = torch.zeros(F, 2, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
W = torch.zeros(F, NODES, 3, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream()
tb = api.TrackBatchC(F, NODES, d["X"].data_ptr(), d["x_offsets"].data_ptr(), d["Y"].data_ptr(), s2.data_ptr(), d["rest"].data_ptr(),
                     d["vis"].data_ptr(), d["vis_offsets"].data_ptr(), d["vis_ext"].data_ptr(), d["vis_ext_offsets"].data_ptr(),
                     None, None, None, None, it.data_ptr(), st.data_ptr(), None)
cb = api.CpdBatchC(F, NODES, d["X"].data_ptr(), d["x_offsets"].data_ptr(), None, d["Y"].data_ptr(), s2.data_ptr(), None, None, None, None,
                   W.data_ptr(), it.data_ptr(), st.data_ptr())
tp = api.TrackParams(max_iter=50, tol=0.0).to_c()
cp = api.CpdParams(max_iter=50, tol=0.0).to_c()


def run():
    d["Y"].copy_(Y0); s2.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if mode == "track":
        ctx.tracking_step_batched_raw(tb, tp, device=True, stream=stream.cuda_stream)
    else:
        ctx.cpd_lle_batched_raw(cb, cp, device=True, stream=stream.cuda_stream)
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


engine = os.environ.get("TDLO_ENGINE", "tq")
for k in ("chunk_points", "truncation", "inflight", "threads"):
    if os.environ.get("TDLO_" + k.upper()):
        ctx.set_option(k, float(os.environ["TDLO_" + k.upper()]))
ctx.set_option("engine", 1 if engine == "tq" else 0)
if engine == "tq":
    clusters = [0]
for c in clusters:
    ctx.set_cluster_size(c)
    run(); run()
    ctx.profile_phases(True)
    ms = min(run() for _ in range(3))
    ph = ctx.profile_phases(False)
    iters = int(it.sum().item()) if mode == "track" else int(it[:, 0].sum().item())
    info = ctx.launch_info()
    print(f"mode={mode} F={F} Nn={NODES} Mp={POINTS} cluster={info['cluster_size']} ctas={info['ctas']} tile={info['tile_points']} occ={info['ctas_per_sm']}: "
          f"{ms:.2f} ms  {iters/ms*1e3:.0f} it/s  {F/ms*1e3:.0f} frames/s")
    if engine == "tq":
        v = list(ph["rank0"].values()) + list(ph["others"].values())
        names = ("wait", "prune", "dmin", "estep", "start_call", "glue", "m_assemble", "m_solve", "m_update", "finish")
        tot = sum(v[:10]) or 1
        print("    phases", {n: f"{x/tot*100:.1f}%" for n, x in zip(names, v[:10])}, f"total {tot/3/1e6:.1f} Mcyc/run over {info['ctas']} CTAs")
        if v[11]:
            print(f"    estep tasks/run {v[10]/3:.0f}  tiles/run {v[11]/3:.0f}  avg window {v[12]/v[11]:.1f} nodes  blocks/tile {v[13]/v[11]:.2f}"
                  f"  cycles/estep-task {v[3]/max(v[10],1):.0f}  cycles/msolve {v[7]*1.0/max(iters*3,1):.0f} assemble {v[6]/max(iters*3,1):.0f} update {v[8]/max(iters*3,1):.0f}")
            print(f"    per estep task: tile loop (avg over warps) {v[14]/max(v[10],1)/(info['threads']//32):.0f} cyc, end barrier+reduce (thread 0) {v[15]/max(v[10],1):.0f} cyc, whole task {v[3]/max(v[10],1):.0f} cyc")
        continue
    for k in ("rank0", "others"):
        tot = sum(ph[k].values()) or 1
        print("   ", k, {n: f"{v/tot*100:.1f}%" for n, v in ph[k].items()}, f"total {tot/3/1e6:.1f} Mcyc/run")
