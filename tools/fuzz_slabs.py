# -*- coding: utf-8 -*-
""" Randomised check of the z-slab decomposition on one GPU (slabs emulated one after the other): random volumes, kernel widths,
pass counts and slab counts (slabs thinner than the halo included) against the undivided run: identical NaN mask,
|fp64 quotient difference| <= 1e-12 x value range; one slab = bit-identical.   python tools/fuzz_slabs.py [cases] [seed] """
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
from fastbarnes import interpolation as fb, distributed as fd
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
bad, done = [], 0
while done < ncases:
    n = int(rng.integers(1, 6))
    size = tuple(int(x) for x in rng.integers(24, 100, 3))
    step = float(rng.choice([0.25, 1.0]))
    sigma = [float(r * step) for r in rng.uniform(1.2, 9.0, 3)]
    if any(2 * fb.get_half_kernel_size_opt(sigma[m], step, n) + 1 >= size[m] for m in range(3)):
        continue
    N = int(rng.integers(200, 6000))
    pts = rng.uniform(-0.03, 1.03, (N, 3)) * (np.asarray(size) - 1) * step
    k = min(N // 3, 100); pts[:k] = pts[k:2 * k]
    val = rng.normal(rng.uniform(-5, 300), rng.uniform(0.5, 20), N)
    vrange = float(val.max() - val.min())
    ref32, ref64 = fb.barnes(pts, val, sigma, [0.0] * 3, step, size, num_iter=n, return_float64=True)
    for nslabs in (1, int(rng.integers(2, 6)), int(rng.integers(6, min(24, size[2]) + 1))):
        got32, got64 = fd.barnes_slabs_emulated(pts, val, sigma, [0.0] * 3, step, size, nslabs, num_iter=n, want_float64=True)
        m = ~np.isnan(ref64)
        nan_ok = bool(np.array_equal(np.isnan(got64), np.isnan(ref64)))
        err = float(np.max(np.abs(got64[m] - ref64[m]))) if m.any() else 0.0
        exact = bool(np.array_equal(got32.view(np.uint32), ref32.view(np.uint32)))
        if (not nan_ok) or err > 1e-12 * vrange or (nslabs == 1 and not exact):
            bad.append({'case': done, 'size': size, 'n': n, 'N': N, 'nslabs': nslabs, 'nan_ok': nan_ok, 'err': err, 'bound': 1e-12 * vrange})
    done += 1
print(json.dumps({'cases_run': done, 'mismatches': bad}))
sys.exit(1 if bad else 0)
