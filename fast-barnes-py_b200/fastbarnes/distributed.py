# -*- coding: utf-8 -*-
"""
Multi-GPU use of the batched path: independent fields (time steps, ensemble members) are
partitioned over the ranks of a torch.distributed process group, one process per GPU.  The path has
no exchange step, so there is NO data-path collective: every rank interpolates its own contiguous
block of fields.  An optional all-gather assembles the full result on every rank (host tensors
over gloo, device tensors over NCCL).

The reference has no counterpart (it is single-threaded); this module only adds plumbing around
`interpolation.barnes_batched`.
"""
import numpy as np

from . import interpolation


def shard_range(nfields, world_size, rank):
    """ Contiguous, balanced block [b0, b1) of `nfields` fields owned by `rank` (first ranks get the remainder). """
    if world_size < 1 or not (0 <= rank < world_size):
        raise RuntimeError('invalid rank / world size: %d / %d' % (rank, world_size))
    base, rem = divmod(int(nfields), int(world_size))
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_samples(sample_offsets, b0, b1):
    """ Sample rows [s0, s1) of the fields [b0, b1) and their offsets rebased to 0. """
    sample_offsets = np.asarray(sample_offsets, dtype=np.int64)
    s0, s1 = int(sample_offsets[b0]), int(sample_offsets[b1])
    return s0, s1, sample_offsets[b0:b1 + 1] - s0


def barnes_batched_sharded(pts, val, sigma, x0, step, size, sample_offsets, method='optimized_convolution',
                           num_iter=4, max_dist=3.5, group=None, gather=True, compute=None):
    """
    Every rank of `group` (default: the world group; without an initialised process group this is a
    plain single-process call) interpolates its block of the B fields described by
    (`pts`, `val`, `sample_offsets`) -- see `interpolation.barnes_batched`.

    gather=True: returns the full float32 array (B,) + size[::-1] on every rank.
    gather=False: returns (b0, local) with `local` the rank's fields [b0, b0 + len(local)).
    `compute` is the per-rank batched interpolation, by default `interpolation.barnes_batched`
    (the CUDA path); tests substitute a checker.
    """
    import torch
    import torch.distributed as dist
    compute = interpolation.barnes_batched if compute is None else compute
    sample_offsets = np.asarray(sample_offsets, dtype=np.int64)
    nfields = len(sample_offsets) - 1
    active = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if active else 1
    rank = dist.get_rank(group) if active else 0
    b0, b1 = shard_range(nfields, world, rank)
    rsize = tuple(int(s) for s in (size if isinstance(size, (list, tuple, np.ndarray)) else (size,)))[::-1]
    if b1 > b0:
        s0, s1, offs = shard_samples(sample_offsets, b0, b1)
        local = compute(pts[s0:s1], val[s0:s1], sigma, x0, step, size, sample_offsets=offs, method=method,
                        num_iter=num_iter, max_dist=max_dist)
        local = np.ascontiguousarray(local, dtype=np.float32).reshape((b1 - b0,) + rsize)
    else:
        local = np.empty((0,) + rsize, dtype=np.float32)
    if not gather:
        return b0, local
    if world == 1:
        return local
    # all ranks need equally shaped tensors: pad every block to the largest one
    per = [shard_range(nfields, world, r) for r in range(world)]
    nmax = max(e - b for b, e in per)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.zeros((nmax,) + rsize, dtype=torch.float32, device=dev)
    if b1 > b0:
        mine[:b1 - b0] = torch.from_numpy(local).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty((nfields,) + rsize, dtype=np.float32)
    for r, (b, e) in enumerate(per):
        if e > b:
            out[b:e] = parts[r][:e - b].cpu().numpy()
    return out
