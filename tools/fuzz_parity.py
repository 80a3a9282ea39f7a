# -*- coding: utf-8 -*-
""" Randomised parity run on a GPU box: random grids (2D / 3D), kernel widths, pass counts, sample counts and field counts
through every kernel family (options sweepq / sweepp), each result compared with the oracle bit for bit.
  python tools/fuzz_parity.py [cases] [seed] """
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
from fastbarnes import interpolation as fb, _lib
from oracle import oracle as orc

L = _lib.lib()
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
mode = sys.argv[3] if len(sys.argv) > 3 else 'small'      # small: few fields; batch: many fields (q kernels with many units); 1d
bad = []
done = 0
for case in range(ncases):
    dim = 1 if mode == '1d' else int(rng.choice([2, 2, 3]))
    n = int(rng.integers(1, 7))
    if dim == 1:
        size = (int(rng.integers(1100, 150000)),)
    elif dim == 2:
        size = tuple(int(x) for x in rng.integers(40, 420, 2))
    else:
        size = tuple(int(x) for x in rng.integers(24, 90, 3))
    step = float(rng.choice([0.1, 0.25, 1.0]))
    ratio = rng.uniform(1.2, 60.0, 1) if dim == 1 else (rng.uniform(1.2, 14.0, dim) if dim == 3 else rng.uniform(1.2, 40.0, dim))
    sigma = [float(r * step) for r in ratio]
    T = [int(fb.get_half_kernel_size_opt(sigma[m], step, n)) for m in range(dim)]
    if any(2 * T[m] + 1 >= size[m] for m in range(dim)):
        continue                                  # the reference refuses these, so do we
    nf = int(rng.choice([1, 1, 2, 5]))
    if mode == 'batch':
        nf = int(rng.choice([16, 40])) if dim == 2 else int(rng.choice([4, 9]))
    N = int(rng.integers(30, 2500))
    ext = (np.asarray(size) - 1) * step
    pts = rng.uniform(-0.03, 1.03, (nf, N, dim)) * ext
    k = min(N // 3, 100)
    pts[:, :k] = pts[:, k:2 * k]                  # repeated locations
    val = rng.normal(rng.uniform(-50, 500), rng.uniform(0.1, 30), (nf, N))
    x0 = [0.0] * dim
    refs = [orc.barnes(pts[i], val[i], sigma, x0, step, size, num_iter=n, nthreads=8) for i in range(nf)]
    for sweepq, sweepp in (((1, 1),) if dim == 1 else ((1, 1), (1, 2), (1, 0), (0, 0))):
        _lib.check(L.fb_set_option(b'sweepq', sweepq))
        _lib.check(L.fb_set_option(b'sweepp', sweepp))
        try:
            out = fb.barnes_batched(pts, val, sigma, x0, step, size, num_iter=n)
        except RuntimeError as e:
            bad.append({'case': case, 'opts': [sweepq, sweepp], 'error': str(e)[:200], 'dim': dim, 'size': size, 'T': T, 'n': n})
            continue
        for i in range(nf):
            if not np.array_equal(out[i].view(np.uint32), refs[i].view(np.uint32)):
                bad.append({'case': case, 'opts': [sweepq, sweepp], 'field': i, 'dim': dim, 'size': size, 'T': T, 'n': n, 'nf': nf, 'N': N,
                            'ndiff': int((out[i].view(np.uint32) != refs[i].view(np.uint32)).sum())})
                break
    done += 1
_lib.check(L.fb_set_option(b'sweepq', 1))
_lib.check(L.fb_set_option(b'sweepp', 1))
print(json.dumps({'cases_run': done, 'option_sets': 4, 'mismatches': bad}))
sys.exit(1 if bad else 0)
