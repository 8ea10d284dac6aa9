"""A C++ host shards tracking_step over two GPUs without Python (tests/cpp/multi_gpu_gather.cpp): device-pointer entry point
per rank + tdlo_all_gather_packed over the host's own NCCL communicator, compared with a single-GPU run of the same batch.
Needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`: profiles/r2_multi_gpu_c.txt)."""
import os
import subprocess

import numpy as np
import pytest

from trackdlo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_gather_symbol_exported():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "trackdlo_b200", "libtrackdlo_b200.so"))
    assert hasattr(lib, "tdlo_all_gather_packed")


@pytest.mark.gpu
def test_cpp_host_shards_over_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from trackdlo_b200 import api
    F, N = 8, 30
    frames = [synth.make_frame(40 + i, n_nodes=N, n_points=1200 + 100 * i, occlusion=0.25 if i % 3 == 1 else 0.0) for i in range(F)]
    X = np.concatenate([f["X"] for f in frames]); xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([len(f["X"]) for f in frames])
    Y = np.stack([f["Y"] for f in frames]); rest = np.stack([f["rest"] for f in frames])
    vis = np.concatenate([f["vis"] for f in frames]).astype(np.int32); vo = np.zeros(F + 1, np.int64); vo[1:] = np.cumsum([len(f["vis"]) for f in frames])
    ext = np.concatenate([f["vis_ext"] for f in frames]).astype(np.int32); eo = np.zeros(F + 1, np.int64); eo[1:] = np.cumsum([len(f["vis_ext"]) for f in frames])
    tp = api.TrackParams(max_iter=10)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fh:
        for a in (np.array([F, N], np.int64), xo, X, Y, rest, vo, vis, eo, ext):
            fh.write(np.ascontiguousarray(a).tobytes())
        fh.write(bytes(tp.to_c()))
    exe = tmp_path / "multi_gpu_gather"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "multi_gpu_gather.cpp"),
                           "-I", os.path.join(ROOT, "include"), "-L", os.path.join(ROOT, "trackdlo_b200"), "-l:libtrackdlo_b200.so", "-lnccl",
                           "-Xlinker", "-rpath", "-Xlinker", os.path.join(ROOT, "trackdlo_b200")])
    r = subprocess.run([str(exe), str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    rec = 3 * N + 4
    got = np.fromfile(out, np.float64).reshape(F, rec)
    ctx = api.Context(max_frames=F, max_nodes=N, max_points_total=len(X))
    ref = ctx.tracking_step_batched(X, xo, Y, np.zeros(F), rest, vis, vo, ext, eo, tp)
    ctx.close()
    gy = got[:, :3 * N].reshape(F, N, 3)
    assert np.abs(gy - ref["Y"]).max() / np.abs(ref["Y"]).max() < 1e-9           # (bit-identical when both contexts pick the same chunk size)
    assert np.allclose(got[:, 3 * N], ref["sigma2"], rtol=1e-9) and np.array_equal(got[:, 3 * N + 3].astype(np.int32), ref["status"])
    assert np.array_equal(got[:, 3 * N + 1:3 * N + 3].astype(np.int32), ref["iters"])
