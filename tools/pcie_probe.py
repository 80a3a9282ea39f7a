#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
pcie_probe.py -- what the host link of the box allows, as the end-to-end arm of bench.py uses it.

  python tools/pcie_probe.py                                         one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py      N GPUs

Per rank: a 704 MiB pinned host buffer (the float32 result of one 64-field sub-batch of the bench workload) copied
device -> host and host -> device with cudaMemcpyAsync, (a) one rank at a time, (b) all ranks at once; ranks bound to the
CPUs next to their GPU first (bench.bind_to_gpu_numa_node), as bench.py does.  Rank 0 prints one JSON object and writes it
to gpurun_out/pcie_probe_<N>gpu.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bench.bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nbytes = 64 * bench.POINTS_PER_FIELD * 4
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(direction, reps=4):
        a, b = (host, d) if direction == 'd2h' else (d, host)
        a.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a.copy_(b, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {}
    for direction in ('d2h', 'h2d'):
        # (a) one rank at a time
        alone = torch.zeros(world, dtype=torch.float64, device=dev)
        for r in range(world):
            barrier()
            if r == rank:
                alone[r] = nbytes / (timed(direction) * 1e-3) / 1e9
        barrier()
        if world > 1:
            dist.all_reduce(alone)
        # (b) all ranks at once
        barrier()
        t = torch.tensor([timed(direction)], dtype=torch.float64, device=dev)
        per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
        per_rank[rank] = nbytes / (float(t[0]) * 1e-3) / 1e9
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(per_rank)
        res[direction] = {'GBps_alone_per_rank': [round(float(x), 2) for x in alone],
                          'GBps_concurrent_per_rank': [round(float(x), 2) for x in per_rank],
                          'GBps_concurrent_aggregate_max_time': round(nbytes * world / (float(t[0]) * 1e-3) / 1e9, 2)}
    if rank == 0:
        out = {'n_gpus': world, 'buffer_bytes': nbytes, 'host_cpus_rank0': numa, 'cpu_count': os.cpu_count(),
               'gpu': torch.cuda.get_device_name(local), **res,
               'e2e_ceiling_grid_points_per_s': res['d2h']['GBps_concurrent_aggregate_max_time'] * 1e9 / 4}
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'pcie_probe_%dgpu.json' % world), 'w') as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
