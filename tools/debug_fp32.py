# -*- coding: utf-8 -*-
""" fp32 working precision against the fp64 path (== reference) on the paper case and random cases. """
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden
from fastbarnes import interpolation as fb

def stats(a, b, name):
    nan_mis = int((np.isnan(a) != np.isnan(b)).sum())
    m = ~np.isnan(a) & ~np.isnan(b)
    d = np.abs(a[m].astype(np.float64) - b[m].astype(np.float64))
    print(name, 'shape', a.shape, 'nan mismatch', nan_mis, 'of', a.size, 'max abs', d.max(), 'rmse', np.sqrt(np.mean(d ** 2)),
          'frac > 1e-3', float((d > 1e-3).mean()), flush=True)

g = load_golden('c1_paper')
step = 1 / 32; x0 = np.asarray([-26 + step, 34.5]); size = (2400, 1200)
for n in (1, 2, 3, 4, 5, 6):
    a = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, num_iter=n, precision='fp32')
    b = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, num_iter=n)
    stats(a, b, 'c1 n=%d' % n)
a = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, num_iter=4, method='convolution', precision='fp32')
b = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, num_iter=4, method='convolution')
stats(a, b, 'c1 plain n=4')
rng = np.random.default_rng(3)
pts = rng.uniform(0, 1, (20000, 2)) * [70, 35]; val = rng.normal(1000, 10, 20000)
for sz, sg in (((1000, 517), 1.0), ((333, 1201), 0.6), ((64, 64), 0.4)):
    a = fb.barnes(pts, val, sg, [0.0, 0.0], 1 / 16, sz, precision='fp32')
    b = fb.barnes(pts, val, sg, [0.0, 0.0], 1 / 16, sz)
    stats(a, b, '2d %s' % (sz,))
p3 = rng.uniform(0, 1, (50000, 3)) * [60, 50, 40]; v3 = rng.normal(0, 1, 50000)
a = fb.barnes(p3, v3, 3.0, [0.0, 0.0, 0.0], 0.5, (121, 101, 81), precision='fp32')
b = fb.barnes(p3, v3, 3.0, [0.0, 0.0, 0.0], 0.5, (121, 101, 81))
stats(a, b, '3d')
