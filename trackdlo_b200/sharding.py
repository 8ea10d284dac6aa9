"""Frame sharding across GPUs: contiguous blocks of frames per rank, no data-path
collective, one all-gather of the tracked node positions (+ sigma2, iters, status) at the end.

Frames are independent problems (each carries its own X, Y, sigma2, visibility), so the EM loop
itself never communicates; a single live sequence does not shard (frame t needs Y of t-1,
trackdlo.cpp:998) -- that case is "replicas only".

The payload of the collective is written by the persistent kernel's epilogue itself
(tdlo_track_batch::packed_results: one contiguous record {Y, sigma2, iters_pre, iters_main, status} per frame),
so a step is exactly two launches per rank: the registration kernel and the NCCL all-gather.
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous [lo, hi) block of ceil(F/G) frames for `rank` (last ranks may get fewer / none)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def record_width(n_nodes):
    """Doubles per frame in tdlo_track_batch::packed_results."""
    return 3 * n_nodes + 4


def alloc_packed(n_frames, n_nodes, world, device):
    """This rank's send buffer: ceil(F/G) records (ranks with fewer frames leave the tail zero)."""
    per = (n_frames + world - 1) // world
    return torch.zeros((per, record_width(n_nodes)), dtype=torch.float64, device=device)


def all_gather_packed(packed_local, gathered=None):
    """ONE collective: every rank ends up with all ranks' records, [world * per, 3 Nn + 4]."""
    world = dist.get_world_size()
    if gathered is None:
        gathered = torch.empty((world * packed_local.shape[0], packed_local.shape[1]), dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(gathered, packed_local)
    return gathered


def unpack(gathered, n_frames, n_nodes):
    """Records -> (Y [F, Nn, 3], sigma2 [F], iters [F, 2] int32, status [F] int32)."""
    g = gathered[:n_frames]
    w = 3 * n_nodes
    return (g[:, :w].reshape(n_frames, n_nodes, 3), g[:, w], g[:, w + 1:w + 3].to(torch.int32), g[:, w + 3].to(torch.int32))
