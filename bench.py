#!/usr/bin/env python
"""bench.py -- EM iterations/s and tracked frames/s of the B200-native TrackDLO registration path.

Workload (BASELINE.json configs[1], "C2"): 64 independent synthetic frames per GPU, Nn=50 nodes,
Mp=20000 points, full trackdlo::tracking_step per frame (trackdlo.cpp:900-999) with max_iter=50 and
tol=0, i.e. exactly 50 EM iterations in the pre-processing registration and 50 in the main one.
A "step" is one batched call over the rank's 64 frames.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line)
  python bench.py --impl reference [...]                         the CPU restatement of the reference
                                                                 (oracle/, all host threads) on the same workload

Keys: `value` = whole-job EM iterations/s with inputs resident in HBM (CUDA events on the launch
stream, max over ranks); `e2e` = same metric through the host-buffer C-ABI call (pinned host
buffers, H2D + kernel + D2H inside the timed region); `roofline` / `fp64` explain the kernel;
`cpu_baseline` is the oracle timed on one host core on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NODES, POINTS, FRAMES_PER_GPU, MAX_ITER = 50, 20000, 64, 50
METRIC, UNIT = "em_iterations_per_sec", "EM iterations/s"
WORKLOAD = ("C2: 64 frames/GPU x tracking_step (pre-proc cpd_lle with LLE + traverse_euclidean + main cpd_lle), "
            "Nn=50, Mp=20000, max_iter=50, tol=0 -> 100 EM iterations per frame")
FP64_PEAK_NOMINAL_TFLOPS = 37.2     # 148 SM x 64 FP64 lanes x 2 x 1.965 GHz; not in MEASURED_PEAKS.json
FP64_PEAK_MEASURED_TFLOPS = 33.0    # dependent-free DFMA loop on this pool's B200 (scripts/micro/fp64pipe.cu,
                                    # profiles/r1_fp64pipe_microbench.txt); DMMA shares the same pipe (37 TF/s alone)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the persistent kernel from the committed ncu --set full
    capture of this same command (profiles/ncu_traffic.json), per launch; None if no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def algorithmic_work(n_nodes, mp_raw, mp_kept, iters):
    """per frame.iteration B_iter = Mp*24 + Nn*32 + Nn*24 + 8 bytes, F_E = 25*Nn*Mp flop;
    per cpd_lle call one-offs (prune + sigma2 init): Mp0*24 bytes, 18*Nn*Mp0 flop."""
    b_iter = mp_kept * 24 + n_nodes * 32 + n_nodes * 24 + 8
    f_iter = 25 * n_nodes * mp_kept
    return iters * b_iter + mp_raw * 24, iters * f_iter + 18 * n_nodes * mp_raw


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(first_frame, n_frames):
    from trackdlo_b200 import synth
    return synth.make_batch(n_frames, first_frame=first_frame, n_nodes=NODES, n_points=POINTS)


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference (oracle/), all host threads
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    per_step = max(1, min(cores, FRAMES_PER_GPU))          # one frame per host thread, up to the C2 batch of 64
    frames = make_workload(0, per_step)["frames"]
    tp = oracle.TrackParams(max_iter=MAX_ITER, tol=0.0)
    oracle.lib()

    def one(f):
        r = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp)
        return int(r["iters"].sum())

    pool = ThreadPoolExecutor(max_workers=per_step)       # ctypes releases the GIL: true parallelism
    for _ in range(min(args.warmup, 1)):
        list(pool.map(one, frames))
    t0 = time.perf_counter(); iters = 0
    for _ in range(args.steps):
        iters += sum(pool.map(one, frames))
    dt = time.perf_counter() - t0
    val = iters / dt
    sample = f"{per_step} frames/step (one per host thread) of the C2 workload, {args.steps} steps, tracking_step via oracle/liboracle.so (g++ -O3)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "frames_per_sec": per_step * args.steps / dt,
            "config": {"workload": WORKLOAD, "nodes": NODES, "points_per_frame": POINTS, "max_iter": MAX_ITER, "tol": 0.0,
                       "frames_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": per_step, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference (Eigen/ROS) is not buildable here; this is the dependency-free C++17 restatement in oracle/ "
                    "(omits the reference's per-call MatrixXd heap traffic, so it is faster than the true reference)"}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_single_core(frames):
    """Oracle on ONE host core (the reference is single-threaded: trackdlo_node.cpp:643), bounded sample."""
    import oracle
    tp = oracle.TrackParams(max_iter=MAX_ITER, tol=0.0)
    oracle.lib()
    t0 = time.perf_counter(); iters = 0; n = 0
    for f in frames:
        r = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp)
        iters += int(r["iters"].sum()); n += 1
        if time.perf_counter() - t0 > 12.0:
            break
    dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": UNIT, "cores": 1, "kind": "port", "frames_per_sec": n / dt,
            "sample": f"{n} frame(s) of the C2 workload (tracking_step, 100 EM iterations each), {dt:.1f} s on one host core, "
                      "oracle/liboracle.so (g++ -std=c++17 -O3, reference flags)"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from trackdlo_b200 import api, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    total_frames = F * world

    wl = make_workload(rank * F, F)
    tp = api.TrackParams(max_iter=MAX_ITER, tol=0.0)
    tpc = tp.to_c()
    ctx = api.Context(max_frames=F, max_nodes=NODES, max_points_total=int(wl["x_offsets"][-1]), device=local_rank)

    # ---------------- pinned host buffers (e2e) and device-resident copies (value)
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    h = {k: pin(wl[k]) for k in ("X", "x_offsets", "Y", "rest", "vis", "vis_offsets", "vis_ext", "vis_ext_offsets")}
    h_Y0 = h["Y"].clone().pin_memory()
    h_s2 = torch.zeros(F, dtype=torch.float64).pin_memory()
    h_guide = torch.zeros(F, NODES, 3, dtype=torch.float64).pin_memory()
    h_pri = torch.zeros(F, 2 * NODES, 4, dtype=torch.float64).pin_memory()
    h_npri = torch.zeros(F, dtype=torch.int32).pin_memory()
    h_iters = torch.zeros(F, 2, dtype=torch.int32).pin_memory()
    h_status = torch.zeros(F, dtype=torch.int32).pin_memory()
    h_state = torch.zeros(F, dtype=torch.int32).pin_memory()
    hb = api.TrackBatchC(F, NODES, h["X"].data_ptr(), h["x_offsets"].data_ptr(), h["Y"].data_ptr(), h_s2.data_ptr(),
                         h["rest"].data_ptr(), h["vis"].data_ptr(), h["vis_offsets"].data_ptr(), h["vis_ext"].data_ptr(),
                         h["vis_ext_offsets"].data_ptr(), None, h_guide.data_ptr(), h_pri.data_ptr(), h_npri.data_ptr(),
                         h_iters.data_ptr(), h_status.data_ptr(), h_state.data_ptr())
    h2d = sum(h[k].numel() * h[k].element_size() for k in h) + h_s2.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in (h["Y"], h_s2, h_guide, h_pri, h_npri, h_iters, h_status, h_state))

    d = {k: h[k].to(dev) for k in h}
    d_Y0 = d["Y"].clone()
    d_s2 = torch.zeros(F, dtype=torch.float64, device=dev)
    d_guide = torch.zeros(F, NODES, 3, dtype=torch.float64, device=dev)
    d_pri = torch.zeros(F, 2 * NODES, 4, dtype=torch.float64, device=dev)
    d_npri = torch.zeros(F, dtype=torch.int32, device=dev)
    d_iters = torch.zeros(F, 2, dtype=torch.int32, device=dev)
    d_status = torch.zeros(F, dtype=torch.int32, device=dev)
    d_state = torch.zeros(F, dtype=torch.int32, device=dev)
    db = api.TrackBatchC(F, NODES, d["X"].data_ptr(), d["x_offsets"].data_ptr(), d["Y"].data_ptr(), d_s2.data_ptr(),
                         d["rest"].data_ptr(), d["vis"].data_ptr(), d["vis_offsets"].data_ptr(), d["vis_ext"].data_ptr(),
                         d["vis_ext_offsets"].data_ptr(), None, d_guide.data_ptr(), d_pri.data_ptr(), d_npri.data_ptr(),
                         d_iters.data_ptr(), d_status.data_ptr(), d_state.data_ptr())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(timed):
        d["Y"].copy_(d_Y0); d_s2.zero_()                 # restore in/out state (outside the timed region)
        flush.zero_()                                    # L2 flush between timed iterations
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        ctx.tracking_step_batched_raw(db, tpc, device=True, stream=stream.cuda_stream)
        if world > 1:                                    # the single all-gather of tracked nodes
            sharding.all_gather_results(d["Y"], d_s2, d_iters[:, 1], d_status, total_frames)
        ev1.record(stream)
        return (ev0, ev1) if timed else None

    def host_step():
        h["Y"].copy_(h_Y0); h_s2.zero_()
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.tracking_step_batched_raw(hb, tpc, device=False)     # H2D + kernel + D2H + sync inside the call
        return time.perf_counter() - t0

    # ---------------- device-resident timing (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                    # nvidia-smi needs ~0.1 s to emit its first sample: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        device_step(False)
    barrier()
    evs = [device_step(True) for _ in range(args.steps)]
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    local_dev_ms = dev_ms
    iters_np = d_iters.cpu().numpy()
    status_np = d_status.cpu().numpy()
    info = ctx.launch_info()
    launches = args.steps * info["launches"]

    # ---------------- end-to-end timing through the host-buffer C-ABI call (`e2e`)
    for _ in range(2):
        host_step()
    barrier()
    e2e_s = sum(host_step() for _ in range(args.steps))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_iters = int(h_iters.numpy().sum())

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        iters_per_step_rank = int(iters_np.sum())
        iters_per_step = iters_per_step_rank * world              # every rank runs the same shape / fixed 100 it per frame
        value = iters_per_step * args.steps / (dev_ms * 1e-3)
        e2e_val = e2e_iters * world * args.steps / (e2e_ms * 1e-3)
        # roofline of the (single) persistent kernel: algorithmic bytes and flops per launch
        xo = wl["x_offsets"]
        alg_b = alg_f = 0.0
        for f in range(F):
            mp0 = int(xo[f + 1] - xo[f])
            for call in range(2):
                nn = len(wl["frames"][f]["vis_ext"]) if call == 0 else NODES
                b, fl = algorithmic_work(nn, mp0, mp0, int(iters_np[f, call]))
                alg_b += b; alg_f += fl
        kern_s = local_dev_ms * 1e-3 / args.steps        # rank 0's own launch (one persistent kernel per step)
        peak, peak_src = load_peaks()
        roof = None
        fp64 = None
        if kern_s:
            ach = alg_b / kern_s / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic_bytes(),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_b,
                    "kernel": "tdlo_tq_kernel<2,256,2> (one persistent task-queue launch per step: prune, E-step chunks, M-steps, traversal)",
                    "note": "the path is 52 flop/B at Nn=50 (SURVEY.md §8d): compute-side bound, HBM fraction is tiny by construction; "
                            "see `fp64` for the FP64 fraction; the E-step is bound by the SMSP issue port (an FP64 instruction holds it ~2.3 cycles, "
                            "profiles/r1_ilp_microbench.txt, DESIGN.md §7)"}
            tf = alg_f / kern_s / 1e12
            fp64 = {"achieved": tf, "peak": FP64_PEAK_MEASURED_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_MEASURED_TFLOPS,
                    "peak_source": "measured DFMA loop (scripts/micro/fp64pipe.cu -> profiles/r1_fp64pipe_microbench.txt)",
                    "frac_of_nominal": tf / FP64_PEAK_NOMINAL_TFLOPS, "nominal_peak": FP64_PEAK_NOMINAL_TFLOPS,
                    "flop_model": "ALGORITHMIC flop, SURVEY.md §8d: 25*Nn*Mp per frame.iteration (+18*Nn*Mp0 per call); the kernel "
                                  "skips affinity entries below exp(-z_cut) (exact zeros / far below one ulp), so executed flop are fewer"}
        cpu = cpu_baseline_single_core(wl["frames"]) if world == 1 else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "frames_per_sec": F * world * args.steps / (dev_ms * 1e-3),
                "config": {"workload": WORKLOAD, "frames_per_gpu": F, "nodes": NODES, "points_per_frame": POINTS,
                           "max_iter": MAX_ITER, "tol": 0.0, "l2": "flushed between timed steps (256 MiB write)",
                           "parallelism": f"frames sharded over {world} GPU(s), one all-gather of results per step" if world > 1 else "single GPU",
                           "launch": info},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms / args.steps, "frames_per_sec": F * world * args.steps / (e2e_ms * 1e-3)},
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roof, "fp64": fp64, "cpu_baseline": cpu,
                "status_mask_or": int(np.bitwise_or.reduce(status_np)),
                "em_iterations_per_step": iters_per_step}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--frames", str(args.frames)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
