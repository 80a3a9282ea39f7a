# -*- coding: utf-8 -*-
""" Where the small-batch kernels (fb_sweepp.cuh) stop paying: F fields of the bench workload (2400x1200, N=50000 per
field, 4 passes), device resident, sweep times with the q kernels (sweepp=0) and the pass-parallel kernels (sweepp=2). """
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
seg, nl = np.zeros(5), np.zeros(1, dtype=np.int64)
res = []
for F in (1, 2, 3, 4, 6, 8, 12, 16):
    pts, val = bench.make_fields(0, F)
    d_p = torch.from_numpy(pts.reshape(F * bench.N_PER_FIELD, 2)).cuda()
    d_v = torch.from_numpy(val.reshape(F * bench.N_PER_FIELD)).cuda()
    plan = fbi.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=F, nsamples=F * bench.N_PER_FIELD,
                            num_iter=bench.NUM_ITER)
    rec = {'fields': F}
    ref = None
    for name, opt in (('q', 0), ('p', 2)):
        _lib.check(L.fb_set_option(b'sweepp', opt))
        for _ in range(3):
            out = plan(d_p, d_v)
        torch.cuda.synchronize()
        L.fb_set_profiling(1)
        acc = np.zeros(5)
        for _ in range(10):
            out = plan(d_p, d_v)
            _lib.check(L.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
            acc += seg
        L.fb_set_profiling(0)
        acc /= 10
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        rec['us_x_' + name] = round(acc[2] * 1e3, 1)
        rec['us_y_' + name] = round(acc[3] * 1e3, 1)
        rec['same_bits'] = bool(np.array_equal(o.view(np.uint32), ref.view(np.uint32)))
    res.append(rec)
    print(json.dumps(rec), flush=True)
_lib.check(L.fb_set_option(b'sweepp', 1))
