import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
from fastbarnes import interpolation as fb
from oracle import oracle as orc
rng = np.random.default_rng(1234)
for lg, n, sigma in [(18, 4, 32.0), (18, 6, 32.0), (20, 4, 32.0), (17, 5, 3.0), (22, 4, 32.0)]:
    L = 2 ** lg; N = L // 64
    pts = rng.uniform(0, L - 1, N); val = rng.normal(0, 1, N)
    t0 = time.perf_counter(); a32, a64 = fb.barnes(pts, val, sigma, 0.0, 1.0, L, num_iter=n, return_float64=True, exact=False); t1 = time.perf_counter()
    o = orc._interpolate_opt_convol(pts.reshape(-1, 1), val.copy(), np.asarray([sigma]), np.zeros(1), np.ones(1), (L,), n, float(np.exp(-3.5**2/2)), stages=True)
    m = ~np.isnan(o['out64'])
    print('L=2^%d n=%d: %.1f ms  nan mask equal %s  max abs diff %.3e  max rel(|q|>1e-3) %.3e  f32 mismatches %d of %d, max ulp %d' % (
        lg, n, (t1 - t0) * 1e3, np.array_equal(np.isnan(a64), np.isnan(o['out64'])), np.max(np.abs(a64[m] - o['out64'][m])),
        np.max((np.abs(a64[m] - o['out64'][m]) / np.abs(o['out64'][m]))[np.abs(o['out64'][m]) > 1e-3]),
        int(np.sum(a32[m] != o['out32'][m])), int(m.sum()),
        int(np.max(np.abs(a32[m].view(np.int32).astype(np.int64) - o['out32'][m].view(np.int32).astype(np.int64))))))
    if lg <= 18:
        e32, e64 = fb.barnes(pts, val, sigma, 0.0, 1.0, L, num_iter=n, return_float64=True, exact=True)
        print('   exact path bit-identical:', np.array_equal(e64[m], o['out64'][m]))
# timing of the C2 shape
L = 2 ** 26; N = 10 ** 6
pts = rng.uniform(0, L - 1, N); val = rng.normal(0, 1, N)
fb.barnes(pts, val, 32.0, 0.0, 1.0, L, num_iter=4)
for n in (4, 6):
    t0 = time.perf_counter(); r = fb.barnes(pts, val, 32.0, 0.0, 1.0, L, num_iter=n); t1 = time.perf_counter()
    print('C2 2^26 N=1e6 n=%d (host API, pageable numpy): %.1f ms, NaN frac %.3f' % (n, (t1 - t0) * 1e3, np.isnan(r).mean()))
