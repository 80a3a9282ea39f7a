import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
from math import exp
from fastbarnes import interpolation as fb
from oracle import oracle as orc
rng = np.random.default_rng(99)
dim, size, sig = 2, (1000, 64), [6.0, 0.3]
N = 700
ext = (np.asarray(size) - 1) * 0.1
pts = rng.uniform(-0.05, 0.8, (N, dim)) * ext
pts[:100] = pts[100:200]
val = rng.normal(3, 20, N)
for n in (3, 4, 6, 7, 8, 10, 12):
    a, a64 = fb.barnes(pts, val, sig, [0.0] * dim, 0.1, size, num_iter=n, return_float64=True)
    st = orc._interpolate_opt_convol(pts, val.copy(), np.asarray(sig), np.zeros(2), np.full(2, 0.1), size, n, exp(-3.5**2/2), stages=True)
    bad = ~((a64 == st['out64']) | (np.isnan(a64) & np.isnan(st['out64'])))
    ys, xs = np.nonzero(bad)
    print('n', n, 'mismatches', bad.sum(), 'x range', (xs.min(), xs.max()) if len(xs) else None, 'y range', (ys.min(), ys.max()) if len(ys) else None,
          'maxrel', np.nanmax(np.abs(a64 - st['out64']) / np.abs(st['out64'])) if bad.sum() else 0)
    # stage check: convolve only
    vg, wg = st['vin'].copy(), st['win'].copy()
    T = fb._get_half_kernel_size_opt(np.asarray(sig), np.full(2, 0.1), n); tv = fb._get_tail_value(np.asarray(sig), np.full(2, 0.1), n)
    fb._convolve_tail_2d(vg, wg, np.asarray(sig), np.full(2, 0.1), size, 2 * T + 1, n, tv, exp(-3.5**2/2))
    badv = ~((vg == st['vg']) | (np.isnan(vg) & np.isnan(st['vg'])))
    print('    T', T, 'convolve stage mismatches', badv.sum(), np.nonzero(badv)[1][:5] if badv.sum() else '')
