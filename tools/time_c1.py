# -*- coding: utf-8 -*-
""" Stage times of the literal paper case (BASELINE configs[0]: ONE 2400x1200 field, N=3490, sigma 1 degree, 4 passes),
device resident: zero fill, min/max + injection, x sweep, y sweep (library's CUDA events) and the whole call. """
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import torch
import bench
from fastbarnes import interpolation as fbi, _lib

L = _lib.lib()
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_paper.npz'))
pts, val = np.ascontiguousarray(g['pts'], dtype=np.float64), np.ascontiguousarray(g['val'], dtype=np.float64)
n = len(val)
plan = fbi.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=1, nsamples=n, num_iter=bench.NUM_ITER)
d_p, d_v = torch.from_numpy(pts).cuda(), torch.from_numpy(val).cuda()
for _ in range(5):
    plan(d_p, d_v)
torch.cuda.synchronize()
seg, acc, nl = np.zeros(5), np.zeros(5), np.zeros(1, dtype=np.int64)
L.fb_set_profiling(1)
for _ in range(20):
    plan(d_p, d_v)
    _lib.check(L.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
    acc += seg
L.fb_set_profiling(0)
acc /= 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    plan(d_p, d_v)
e1.record()
torch.cuda.synchronize()
print(json.dumps({'N': n, 'us_zero_fill': acc[0] * 1e3, 'us_minmax_inject': acc[1] * 1e3, 'us_sweep_x': acc[2] * 1e3,
                  'us_sweep_y': acc[3] * 1e3, 'us_whole_call_back_to_back': e0.elapsed_time(e1) * 1e3 / 50,
                  'kernel_launches_per_call': int(nl[0])}))
