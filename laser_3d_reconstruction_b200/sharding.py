"""Frame-wise sharding over the GPUs of one box.

Frames are independent, so rank r simply takes frames r::world -- no collective on the hot path.
The only exchange is the final gather of the per-frame point clouds to rank 0: one all_gather of
the per-rank row counts, then a padded gather of the rows (NCCL over NVLink on GPUs; gloo in the
CPU tests).  torch.distributed is plumbing here, nothing more.
"""
import numpy as np


def shard_frames(nframes, rank, world):
    """Indices of the frames rank `rank` of `world` processes (r::world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return list(range(rank, nframes, world))


def pack_clouds(frame_ids, clouds):
    """[(frame_id, Nx3 array)] -> one (sum N) x 4 float64 table with the frame id in column 0."""
    rows = [np.zeros((0, 4), np.float64)]
    for f, c in zip(frame_ids, clouds):
        c = np.asarray(c, np.float64).reshape(-1, 3)
        rows.append(np.concatenate([np.full((len(c), 1), float(f)), c], axis=1))
    return np.concatenate(rows, axis=0)


def unpack_clouds(table, nframes):
    """Inverse of pack_clouds over all ranks: list of Nx3 arrays indexed by frame id."""
    table = np.asarray(table, np.float64).reshape(-1, 4)
    ids = table[:, 0].astype(np.int64)
    order = np.argsort(ids, kind="stable")
    table, ids = table[order], ids[order]
    bounds = np.searchsorted(ids, np.arange(nframes + 1))
    return [table[bounds[f]:bounds[f + 1], 1:] for f in range(nframes)]


def gather_point_clouds(table, device=None, group=None, dst=0, to_host=True):
    """Gather each rank's packed (N_r x 4) table to rank `dst`.  Returns the concatenated table on
    dst and None elsewhere.  Works with any initialised torch.distributed backend; `device` is the
    torch device the backend needs its tensors on (cuda:<local_rank> for nccl, cpu for gloo).
    `to_host=False` leaves the gathered table on dst's device (a torch tensor): the exchange then costs the
    collective only, no device-to-host copy of the whole job's point clouds per step."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if isinstance(table, torch.Tensor):
            return table.detach().cpu().numpy().reshape(-1, 4)
        return np.asarray(table, np.float64).reshape(-1, 4)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if isinstance(table, torch.Tensor):  # already resident on the backend's device (packed by the C ABI)
        t = table.reshape(-1, 4)
    else:
        t = torch.from_numpy(np.ascontiguousarray(table, np.float64).reshape(-1, 4)).to(dev)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    padded = torch.empty((nmax, 4), dtype=torch.float64, device=dev)
    padded[:t.shape[0]] = t
    padded[t.shape[0]:] = 0
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        if not to_host:
            return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
        return np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)
    dist.gather(padded, None, dst=dst, group=group)
    return None
