"""
3D z-slab run under torchrun (one process per GPU, NCCL halo exchange):
  torchrun --nproc-per-node N tools/slab_multi_gpu.py [W H D N sigma]
Checks the assembled volume against the single-GPU result (rank 0) and reports timings.
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
import torch
import torch.distributed as dist
from fastbarnes import interpolation as fb, distributed as fd, _lib

rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
_lib.check(_lib.lib().fb_set_device(local))
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
args = sys.argv[1:]
W, H, D, N, sigma = (int(args[0]), int(args[1]), int(args[2]), int(args[3]), float(args[4])) if len(args) >= 5 else (512, 512, 256, 1250000, 8.0)
rng = np.random.default_rng(1235)
pts = rng.uniform(0, 1, (N, 3)) * [W - 1, H - 1, D - 1]
val = rng.normal(0, 1, N)
dp = torch.from_numpy(pts).to(dev); dv = torch.from_numpy(val).to(dev)
slab = fd.BarnesSlab3D(sigma, [0.0] * 3, 1.0, (W, H, D), N, num_iter=4, want_float64=True)
out = slab(dp, dv)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
def timed(k=5):
    ts = []
    for _ in range(k):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        slab.phase1(dp, dv); torch.cuda.synchronize(); t1 = time.perf_counter()
        slab.exchange(); torch.cuda.synchronize(); t2 = time.perf_counter()
        slab.phase2(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t3 = time.perf_counter()
        ts.append((t3 - t0, t1 - t0, t2 - t1, t3 - t2))
    return min(ts)
tot, p1, ex, p2 = timed()
# gather the own planes on rank 0 and compare with the single-GPU run
res = {'world': world, 'grid': [W, H, D], 'N': N, 'sigma': sigma, 'halo_planes': slab.halo, 'ms_total': tot * 1e3, 'ms_phase1': p1 * 1e3,
       'ms_exchange': ex * 1e3, 'ms_phase2': p2 * 1e3, 'grid_points_per_s': W * H * D / tot,
       'halo_bytes_per_side': slab.halo * W * H * 16}
if world > 1:
    parts = [torch.empty((fd.shard_range(D, world, r)[1] - fd.shard_range(D, world, r)[0], H, W), dtype=torch.float64, device=dev)
             for r in range(world)] if rank == 0 else None
    # shapes are equal when D % world == 0
    dist.gather(slab.out64, parts, dst=0)
if rank == 0:
    full64 = torch.cat(parts, 0).cpu().numpy() if world > 1 else slab.out64.cpu().numpy()
    plan = fb.BarnesDevice(3, sigma, [0.0] * 3, 1.0, (W, H, D), nfields=1, nsamples=N, num_iter=4, want_float64=True)
    plan(dp, dv); torch.cuda.synchronize()
    t0 = time.perf_counter(); plan(dp, dv); torch.cuda.synchronize(); t_single = time.perf_counter() - t0
    ref64 = plan.out64[0].cpu().numpy()
    m = ~np.isnan(ref64)
    res['nan_mask_equal'] = bool(np.array_equal(np.isnan(full64), np.isnan(ref64)))
    res['max_rel_diff_fp64'] = float(np.max(np.abs(full64[m] - ref64[m]) / np.maximum(np.abs(ref64[m]), 1e-300)))
    res['max_abs_diff_fp64'] = float(np.max(np.abs(full64[m] - ref64[m])))
    res['ms_single_gpu_device_resident'] = t_single * 1e3
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'slab_%dgpu.json' % world), 'w'), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
