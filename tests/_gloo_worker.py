"""world_size-2 gloo worker: exercises avatar_b200.shard (the multi-rank plumbing bench.py uses) on CPU"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatar_b200 import shard  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_frames, nx = 10, 109
lo, hi = shard.frame_range(n_frames, rank, world)
local = np.zeros((hi - lo, nx))
local[:, 0] = np.arange(lo, hi)          # "fitted parameters" tagged with the global frame id
full = shard.gather_params(local, n_frames, rank, world)
t = shard.max_over_ranks(float(rank + 1))
assert t == float(world)
if rank == 0:
    np.save(sys.argv[1], full)
dist.destroy_process_group()
