"""Frame sharding across GPUs: contiguous blocks of frames per rank, no data-path
collective, one all-gather of the tracked node positions (+ sigma2, iters, status) at the end.

Frames are independent problems (each carries its own X, Y, sigma2, visibility), so the EM loop
itself never communicates; a single live sequence does not shard (frame t needs Y of t-1,
trackdlo.cpp:998) -- that case is "replicas only".
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous [lo, hi) block of ceil(F/G) frames for `rank` (last ranks may get fewer / none)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def all_gather_results(Y_local, sigma2_local, iters_local, status_local, n_frames):
    """Every rank ends up with the full [F, Nn, 3] trajectory block.  Inputs are this rank's shard
    (tensors on the rank's device, or CPU tensors under gloo).  Shards are padded to ceil(F/G)."""
    world = dist.get_world_size()
    per = (n_frames + world - 1) // world
    Nn = Y_local.shape[1]
    dev = Y_local.device

    def pad(t, shape, dtype):
        out = torch.zeros(shape, dtype=dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    # one flat fp64 payload per rank: Y | sigma2 | iters | status  -> a single collective
    width = Nn * 3 + 3
    payload = torch.zeros((per, width), dtype=torch.float64, device=dev)
    n = Y_local.shape[0]
    payload[:n, : Nn * 3] = Y_local.reshape(n, Nn * 3)
    payload[:n, Nn * 3] = sigma2_local
    payload[:n, Nn * 3 + 1] = iters_local.to(torch.float64)
    payload[:n, Nn * 3 + 2] = status_local.to(torch.float64)
    gathered = torch.empty((world * per, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, payload)
    gathered = gathered[:n_frames]
    Y = gathered[:, : Nn * 3].reshape(n_frames, Nn, 3)
    return Y, gathered[:, Nn * 3], gathered[:, Nn * 3 + 1].to(torch.int32), gathered[:, Nn * 3 + 2].to(torch.int32)
