# -*- coding: utf-8 -*-
"""
CPU tests of the N>1 host logic (gloo, world_size 2): field partitioning and result assembly of
`fastbarnes.distributed`.  The per-rank compute is substituted by the CPU oracle (as the checker),
because these tests run without a GPU; the CUDA per-rank path is covered by the -m gpu tests.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, bits_equal
from fastbarnes import distributed


def test_shard_range_partitions_exactly():
    for nfields in (0, 1, 5, 8, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            blocks = [distributed.shard_range(nfields, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nfields
            for (a0, a1), (b0, b1) in zip(blocks[:-1], blocks[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(RuntimeError):
        distributed.shard_range(4, 2, 2)


def test_shard_samples_rebases_offsets():
    offs = np.asarray([0, 3, 3, 10, 14])
    s0, s1, local = distributed.shard_samples(offs, 1, 4)
    assert (s0, s1) == (3, 14) and list(local) == [0, 0, 7, 11]


def _oracle_batched(pts, val, sigma, x0, step, size, sample_offsets=None, method='optimized_convolution', num_iter=4,
                    max_dist=3.5):
    from oracle import oracle as orc
    out = []
    for b in range(len(sample_offsets) - 1):
        s0, s1 = sample_offsets[b], sample_offsets[b + 1]
        out.append(orc.barnes(pts[s0:s1], val[s0:s1], sigma, x0, step, size, method=method, num_iter=num_iter,
                              max_dist=max_dist))
    return np.stack(out)


def _make_problem():
    rng = np.random.default_rng(42)
    counts = [40, 75, 20, 64, 33]
    offs = np.concatenate([[0], np.cumsum(counts)])
    pts = rng.uniform(0, 1, (offs[-1], 2)) * np.asarray([7.9, 5.9])
    val = rng.normal(0, 1, offs[-1])
    return pts, val, offs


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fastbarnes import distributed as fd
        pts, val, offs = _make_problem()
        full = fd.barnes_batched_sharded(pts, val, 0.6, [0.0, 0.0], 0.1, (80, 60), offs, num_iter=3,
                                         compute=_oracle_batched)
        b0, local = fd.barnes_batched_sharded(pts, val, 0.6, [0.0, 0.0], 0.1, (80, 60), offs, num_iter=3,
                                              compute=_oracle_batched, gather=False)
        np.savez(os.path.join(tmpdir, 'rank%d.npz' % rank), full=full, b0=b0, local=local)
    finally:
        dist.destroy_process_group()


def test_sharded_batched_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pts, val, offs = _make_problem()
    want = _oracle_batched(pts, val, 0.6, [0.0, 0.0], 0.1, (80, 60), sample_offsets=offs, num_iter=3)
    seen = 0
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % rank))
        assert bits_equal(z['full'], want)                      # every rank holds the assembled result
        b0 = int(z['b0'])
        assert bits_equal(z['local'], want[b0:b0 + len(z['local'])])
        seen += len(z['local'])
    assert seen == 5


def test_single_process_is_plain_call():
    pts, val, offs = _make_problem()
    got = distributed.barnes_batched_sharded(pts, val, 0.6, [0.0, 0.0], 0.1, (80, 60), offs, num_iter=3,
                                             compute=_oracle_batched)
    assert bits_equal(got, _oracle_batched(pts, val, 0.6, [0.0, 0.0], 0.1, (80, 60), sample_offsets=offs, num_iter=3))


def test_slab_transfer_plan_covers_every_halo_plane():
    """ z-slab halo exchange plan (fastbarnes.distributed.slab_transfers): for every rank the planes of its extended window
    that it does not own arrive exactly once, from their owner; sends and receives of the two ends of a transfer agree;
    slabs thinner than the halo take planes from several ranks. """
    from fastbarnes.distributed import slab_transfers, shard_range
    for nplanes, world, halo in ((512, 8, 28), (512, 2, 28), (120, 16, 12), (40, 8, 28), (64, 4, 1), (30, 3, 0)):
        plans = [slab_transfers(nplanes, world, r, halo) for r in range(world)]
        for r in range(world):
            z0, z1 = shard_range(nplanes, world, r)
            e0, e1 = max(0, z0 - halo), min(nplanes, z1 + halo)
            got = []
            for q, send, recv in plans[r]:
                if recv:
                    q0, q1 = shard_range(nplanes, world, q)
                    assert q0 <= recv[0] < recv[1] <= q1            # the sender owns what it sends
                    got.extend(range(recv[0], recv[1]))
                    # the peer's plan holds the matching send
                    assert any(p == r and s == recv for p, s, _ in plans[q])
                if send:
                    assert z0 <= send[0] < send[1] <= z1
                    assert any(p == r and rc == send for p, _, rc in plans[q])
            want = [z for z in range(e0, e1) if not (z0 <= z < z1)]
            assert sorted(got) == want, (nplanes, world, halo, r)
