# -*- coding: utf-8 -*-
""" BASELINE configs[1] (1D, N = L/64 uniform samples, sigma 32): stage times of the exact path (line kernel), the old
lane-pair walk and the segmented option, device resident.  Usage: python tools/time_1d.py [log2_L] """
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import torch
from fastbarnes import interpolation as fbi, _lib

L_ = _lib.lib()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
L = 2 ** lg
N = L // 64
rng = np.random.default_rng(1234)
pts = rng.uniform(0, L - 1, N)
val = rng.normal(0, 1, N)
d_p, d_v = torch.from_numpy(pts).cuda(), torch.from_numpy(val).cuda()
seg = np.zeros(5)
nl = np.zeros(1, dtype=np.int64)
out = {}
for n in (4, 6):
    plan = fbi.BarnesDevice(1, 32.0, 0.0, 1.0, L, nfields=1, nsamples=N, num_iter=n)
    for name, opt in (('line_kernel', 1), ('lane_pair_walk', 0)):
        if opt == 0 and lg > 22:
            continue
        _lib.check(L_.fb_set_option(b'line1d', opt))
        plan(d_p, d_v)
        torch.cuda.synchronize()
        L_.fb_set_profiling(1)
        plan(d_p, d_v)
        _lib.check(L_.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
        L_.fb_set_profiling(0)
        out['n%d_%s' % (n, name)] = {'ms_zero_fill': round(seg[0], 3), 'ms_inject': round(seg[1], 3), 'ms_line': round(seg[2], 3),
                                     'ns_per_point': round(seg[2] * 1e6 / L, 2)}
    _lib.check(L_.fb_set_option(b'line1d', 1))
print(json.dumps({'L': L, 'N': N, 'sigma_over_step': 32.0, 'results': out}, indent=1))
