# -*- coding: utf-8 -*-
""" One fp32 batch of the bench workload (for ncu captures of fb_sweep32_kernel). """
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import torch
import bench
from fastbarnes import interpolation as fbi
F = int(os.environ.get('FIELDS', '64'))
dev = torch.device('cuda', 0)
pts_h, val_h = bench.make_fields(0, F)
d_pts = torch.from_numpy(pts_h.reshape(F * bench.N_PER_FIELD, 2)).to(dev)
d_val = torch.from_numpy(val_h.reshape(F * bench.N_PER_FIELD)).to(dev)
plan = fbi.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=F, nsamples=F * bench.N_PER_FIELD,
                        num_iter=bench.NUM_ITER, device=dev, precision=os.environ.get('PRECISION', 'fp32'))
for _ in range(int(os.environ.get('REPS', '3'))):
    plan(d_pts, d_val)
torch.cuda.synchronize()
print('done')
