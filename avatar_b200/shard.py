"""Multi-GPU plumbing: independent frames shard one contiguous block per rank (no data-path collective);
a single all_gather of the fitted parameters at the end (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def frame_range(n_frames, rank, world):
    """contiguous, balanced block of frames for this rank"""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_params(local, n_frames, rank, world, device=None):
    """all_gather of [frames_per_rank, nx] float64 (padded to equal size) -> [n_frames, nx] on every rank"""
    import torch
    import torch.distributed as dist
    nx = local.shape[1]
    per = (n_frames + world - 1) // world
    buf = torch.zeros((per, nx), dtype=torch.float64, device=device)
    buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    if world > 1:
        dist.all_gather(out, buf)
    else:
        out = [buf]
    rows = []
    for r in range(world):
        lo, hi = frame_range(n_frames, r, world)
        rows.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(rows, axis=0)


def comm_unique_id():
    """a fresh 128-byte NCCL id (rank 0 makes it, the others get it through whatever the launcher offers)"""
    import ctypes as C
    from ._lib import lib, check
    buf = (C.c_uint8 * 128)()
    check(lib.avb_comm_unique_id(buf))
    return bytes(buf)


class ParamGatherer:
    """The per-step gather of the fitted parameters with everything allocated ONCE: a pinned host staging buffer, one
    device send buffer, one device receive buffer for all_gather_into_tensor and one pinned host copy of the result
    (the per-step tensor allocations and the eight .cpu() calls cost more than the collective)."""

    def __init__(self, per_rank, nx, rank, world, device, ranges):
        import torch
        self.per, self.nx, self.rank, self.world, self.ranges = per_rank, nx, rank, world, ranges
        on_gpu = device is not None and torch.device(device).type == "cuda"
        self.h_in = torch.zeros((per_rank, nx), dtype=torch.float64, pin_memory=on_gpu)
        self.d_in = torch.zeros((per_rank, nx), dtype=torch.float64, device=device)
        self.d_out = torch.zeros((world * per_rank, nx), dtype=torch.float64, device=device)
        self.h_out = torch.zeros((world * per_rank, nx), dtype=torch.float64, pin_memory=on_gpu)

    def gather(self, local):
        import torch
        import torch.distributed as dist
        n = local.shape[0]
        self.h_in[:n] = torch.from_numpy(np.ascontiguousarray(local))
        self.d_in.copy_(self.h_in, non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_out, self.d_in)
        else:
            self.d_out.copy_(self.d_in)
        self.h_out.copy_(self.d_out, non_blocking=True)
        if self.d_out.is_cuda:
            torch.cuda.current_stream(self.d_out.device).synchronize()
        full = self.h_out.numpy()
        return np.concatenate([full[r * self.per:r * self.per + (hi - lo)] for r, (lo, hi) in enumerate(self.ranges)], axis=0)


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
