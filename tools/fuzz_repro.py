""" Repeats one configuration of tools/fuzz_parity.py many times and counts mismatches against the oracle.
  python tools/fuzz_repro.py <case> <seed> [repeats] """
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
from fastbarnes import interpolation as fb, _lib
from oracle import oracle as orc
L = _lib.lib()
want, seed = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
rng = np.random.default_rng(seed)
for case in range(want + 1):
    dim = int(rng.choice([2, 2, 3])); n = int(rng.integers(1, 7))
    size = tuple(int(x) for x in rng.integers(40, 420, 2)) if dim == 2 else tuple(int(x) for x in rng.integers(24, 90, 3))
    step = float(rng.choice([0.1, 0.25, 1.0]))
    ratio = rng.uniform(1.2, 14.0, dim) if dim == 3 else rng.uniform(1.2, 40.0, dim)
    sigma = [float(r * step) for r in ratio]
    T = [int(fb.get_half_kernel_size_opt(sigma[m], step, n)) for m in range(dim)]
    if any(2 * T[m] + 1 >= size[m] for m in range(dim)):
        continue
    nf = int(rng.choice([1, 1, 2, 5])); N = int(rng.integers(30, 2500))
    ext = (np.asarray(size) - 1) * step
    pts = rng.uniform(-0.03, 1.03, (nf, N, dim)) * ext
    k = min(N // 3, 100); pts[:, :k] = pts[:, k:2 * k]
    val = rng.normal(rng.uniform(-50, 500), rng.uniform(0.1, 30), (nf, N))
    if case != want:
        continue
    x0 = [0.0] * dim
    refs = [orc.barnes(pts[i], val[i], sigma, x0, step, size, num_iter=n, nthreads=8) for i in range(nf)]
    res = {}
    for opts in ((1, 1), (1, 0), (1, 2), (0, 0)):
        _lib.check(L.fb_set_option(b'sweepq', opts[0])); _lib.check(L.fb_set_option(b'sweepp', opts[1]))
        badruns = 0; fields = set()
        for r in range(reps):
            out = fb.barnes_batched(pts, val, sigma, x0, step, size, num_iter=n)
            for i in range(nf):
                if not np.array_equal(out[i].view(np.uint32), refs[i].view(np.uint32)):
                    badruns += 1; fields.add(i); break
        res[str(opts)] = {'bad_runs': badruns, 'of': reps, 'fields': sorted(fields)}
    print(json.dumps({'case': want, 'dim': dim, 'size': size, 'T': T, 'n': n, 'nf': nf, 'N': N, 'results': res}))
