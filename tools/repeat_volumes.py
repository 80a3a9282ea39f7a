""" Determinism of mid-size 3D volumes (q kernels: transposing x, in-place y, finalising z): 60 runs each must reproduce the first one bit for bit. """
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np, torch
from fastbarnes import interpolation as fb
rng = np.random.default_rng(3)
res = []
for size, sig, n, nf in (((256, 256, 128), 6.0, 4, 1), ((200, 150, 90), 3.5, 2, 3), ((300, 200, 60), 9.0, 3, 2), ((128, 128, 128), 4.0, 1, 4)):
    N = 200000
    pts = rng.uniform(0, 1, (nf * N, 3)) * (np.asarray(size) - 1)
    val = rng.normal(0, 1, nf * N)
    plan = fb.BarnesDevice(3, sig, [0.0] * 3, 1.0, size, nfields=nf, nsamples=nf * N, num_iter=n)
    dp, dv = torch.from_numpy(pts).cuda(), torch.from_numpy(val).cuda()
    first = plan(dp, dv).clone()
    bad = sum(0 if torch.equal(plan(dp, dv).view(torch.int32), first.view(torch.int32)) else 1 for _ in range(60))
    res.append({'size': size, 'n': n, 'fields': nf, 'runs': 60, 'differing': bad})
print(json.dumps(res))
