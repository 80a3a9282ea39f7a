import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
from fastbarnes import interpolation as fb
from oracle import oracle as orc
rng = np.random.default_rng(1234)
lg, n, sigma = 18, 4, 32.0
L = 2 ** lg; N = L // 64
pts = rng.uniform(0, L - 1, N); val = rng.normal(0, 1, N)
a32, a64 = fb.barnes(pts, val, sigma, 0.0, 1.0, L, num_iter=n, return_float64=True, exact=False)
o = orc._interpolate_opt_convol(pts.reshape(-1, 1), val.copy(), np.asarray([sigma]), np.zeros(1), np.ones(1), (L,), n, float(np.exp(-3.5**2/2)), stages=True)
m = ~np.isnan(o['out64'])
d = np.abs(a64 - o['out64']); d[~m] = 0
bad = np.nonzero(d > 1e-9)[0]
print('bad count', len(bad), 'seg 448')
print('positions', bad[:40])
print('pos mod 448', (bad % 448)[:40])
print('diffs', d[bad][:10], 'q', o['out64'][bad][:10], 'w', o['wg'][bad][:10], 'v', o['vg'][bad][:10])
# relative error in w and v implied? compute our v,w not available; look at neighbours
i = bad[0]
print('around', i, a64[i-3:i+4], o['out64'][i-3:i+4], o['wg'][i-3:i+4])
