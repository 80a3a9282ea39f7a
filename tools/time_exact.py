# -*- coding: utf-8 -*-
""" Times the exact-sum methods at the paper's full grid (2400x1200, N=3490) through the public API and
reports the RMSE of the optimized convolution against them (the reference's accuracy yardstick,
demo/rmse.py convention).  Run on the GPU box:  python tools/time_exact.py """
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden
from fastbarnes import interpolation, interpolationS2

g = load_golden('c1_paper')
pts, val = g['pts'], g['val']
step = 1.0 / 32
x0 = np.asarray([-26.0 + step, 34.5])
size = (2400, 1200)
res = {}
for name, fn, kw in (('naive', interpolation.barnes, dict(method='naive')),
                     ('radius', interpolation.barnes, dict(method='radius')),
                     ('naive_S2', interpolationS2.barnes_S2, dict(method='naive_S2'))):
    fn(pts, val, 1.0, x0, step, size, **kw)
    t0 = time.perf_counter()
    out = fn(pts, val, 1.0, x0, step, size, **kw)
    res[name + '_s'] = time.perf_counter() - t0
    res[name] = out
conv = {n: interpolation.barnes(pts, val, 1.0, x0, step, size, num_iter=n) for n in (2, 4, 6)}
s2 = interpolationS2.barnes_S2(pts, val, 1.0, x0, step, size, method='optimized_convolution_S2', num_iter=4)
out = {k: v for k, v in res.items() if k.endswith('_s')}
for n, c in conv.items():
    m = ~np.isnan(c)
    out['rmse_conv_n%d_vs_naive' % n] = float(np.sqrt(np.mean((c[m] - res['naive'][m]) ** 2)))
m = ~np.isnan(s2)
out['rmse_convS2_n4_vs_naive_S2'] = float(np.sqrt(np.mean((s2[m] - res['naive_S2'][m]) ** 2)))
out['pairs'] = int(len(val)) * size[0] * size[1]
print(json.dumps(out))
