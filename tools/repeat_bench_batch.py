""" Runs the 64-field bench batch (2400x1200, N=50000) many times on one GPU and checks that every run reproduces the first one
bit for bit (all 64 fields), and that fields 0, 31, 63 equal the oracle.   python tools/repeat_bench_batch.py [runs] """
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np, torch
import bench
from fastbarnes import interpolation as fb
from oracle import oracle as orc
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
F, N = bench.SUB_FIELDS, bench.N_PER_FIELD
pts, val = bench.make_fields(0, F)
plan = fb.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=F, nsamples=F * N, num_iter=bench.NUM_ITER)
dp, dv = torch.from_numpy(pts.reshape(F * N, 2)).cuda(), torch.from_numpy(val.reshape(F * N)).cuda()
first = plan(dp, dv).clone()
bad = 0
for r in range(runs):
    out = plan(dp, dv)
    if not torch.equal(out.view(torch.int32), first.view(torch.int32)):
        bad += 1
ok = True
for i in (0, 31, F - 1):
    ref = orc.barnes(pts[i], val[i], bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, num_iter=bench.NUM_ITER, nthreads=8)
    ok = ok and bool(np.array_equal(first[i].cpu().numpy().view(np.uint32), ref.view(np.uint32)))
print(json.dumps({'runs': runs, 'runs_differing_from_the_first': bad, 'oracle_fields_equal': ok}))
