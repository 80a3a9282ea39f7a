"""Times the BASELINE.json configurations through the public API on one GPU (wall clock, best of k)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
import torch
from fastbarnes import interpolation as fb, interpolationS2 as fbS2, _lib
from conftest_shim import load_golden

def best(f, k=5):
    f(); ts = []
    for _ in range(k):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return min(ts)

res = {}
g = load_golden('c1_paper')
pts, val = g['pts'], g['val']
step = 1 / 32; x0 = np.asarray([-26 + step, 34.5]); size = (2400, 1200)
res['C1 2D paper 2400x1200 N=3490 n=4 (host API, numpy in/out)'] = best(lambda: fb.barnes(pts, val, 1.0, x0, step, size, num_iter=4))
plan = fb.BarnesDevice(2, 1.0, x0, step, size, nfields=1, nsamples=len(val), num_iter=4)
dp = torch.from_numpy(pts).cuda(); dv = torch.from_numpy(val).cuda()
def dev():
    plan(dp, dv); torch.cuda.synchronize()
res['C1 device-resident (BarnesDevice, sync)'] = best(dev, 20)
for r in (32, 64):
    step = 1 / r; x0 = np.asarray([-26 + step, 34.5]); size = (int(75 / step), int(37.5 / step))
    res['C4 S2 res %d %dx%d N=3490 n=4 (host API)' % (r, size[0], size[1])] = best(lambda: fbS2.barnes_S2(pts, val, 1.0, x0, step, size, method='optimized_convolution_S2', num_iter=4), 3)
rng = np.random.default_rng(1235)
for (W, H, D, N) in [(256, 256, 128, 156250), (512, 512, 256, 1250000)]:
    p3 = rng.uniform(0, 1, (N, 3)) * [W - 1, H - 1, D - 1]; v3 = rng.normal(0, 1, N)
    sig = 4.0 if W == 256 else 8.0
    res['C3-like 3D %dx%dx%d N=%d sigma=%g n=4 (host API)' % (W, H, D, N, sig)] = best(lambda: fb.barnes(p3, v3, sig, [0.0] * 3, 1.0, (W, H, D), num_iter=4), 2)
rng = np.random.default_rng(1234)
for lg in (20, 22):
    L = 2 ** lg; N = L // 64
    p1 = rng.uniform(0, L - 1, N); v1 = rng.normal(0, 1, N)
    res['C2-like 1D 2^%d N=%d sigma=32 n=4 (host API, exact sequential)' % (lg, N)] = best(lambda: fb.barnes(p1, v1, 32.0, 0.0, 1.0, L, num_iter=4), 1)
for k, v in res.items():
    print('%-75s %10.3f ms' % (k, v * 1e3))
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'config_times.json'), 'w'), indent=1)
