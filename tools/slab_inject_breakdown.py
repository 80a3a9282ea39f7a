""" Kernel times of the injection of ONE slab of the C3 volume (slab 3 of 8: 64 planes) on one GPU:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/slab_inject_breakdown.py """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
import torch
from fastbarnes import distributed as fd
W, H, D, N, sigma = 1024, 1024, 512, 10_000_000, 8.0
rng = np.random.default_rng(1235)
pts = rng.uniform(0, 1, (N, 3)) * [W - 1, H - 1, D - 1]
val = rng.normal(0, 1, N)
dp = torch.from_numpy(pts).cuda(); dv = torch.from_numpy(val).cuda()
nsl = int(sys.argv[1]) if len(sys.argv) > 1 else 8
slab = fd.BarnesSlab3D(sigma, [0.0] * 3, 1.0, (W, H, D), N, num_iter=4, nslabs=nsl, slab=nsl // 2 - 1)
for _ in range(2):
    slab.inject(dp, dv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); slab.inject(dp, dv); e1.record(); torch.cuda.synchronize()
print('inject ms', e0.elapsed_time(e1), 'planes', slab.zc)
