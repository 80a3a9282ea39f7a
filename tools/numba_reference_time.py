# -*- coding: utf-8 -*-
""" Times the UNMODIFIED reference (Numba, single thread) in the build container, for orientation next to the C port that
bench.py times on the GPU box (the reference's sources are not in this repository and /root/reference does not exist on
the GPU box).  Convention of the reference's own timing scripts: warm-up call (JIT) excluded, best of 5
(demo/timing1_table1_figure4.py:60, 98-116).   Usage: python tools/numba_reference_time.py > profiles/r2_numba_reference_cpu.json """
import json
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbarnes import interpolation as ref          # noqa: E402  (the reference package)
import bench                                          # noqa: E402

out = {'what': 'fastbarnes.interpolation.barnes(method=optimized_convolution, num_iter=4), reference v2.0.0 under Numba, 1 thread',
       'where': 'build container (no GPU): %s, %d logical CPUs' % (platform.processor() or platform.machine(), os.cpu_count())}
cases = {}
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'c1_paper.npz'))
cases['c1_paper_N3490'] = (np.asarray(g['pts']), np.asarray(g['val']))
p, v = bench.make_fields(0, 1)
cases['c5_field_N50000'] = (p[0], v[0])
for name, (pts, val) in cases.items():
    ref.barnes(pts, val, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, num_iter=bench.NUM_ITER)     # JIT
    best = None
    for _ in range(5):
        t0 = time.perf_counter()
        ref.barnes(pts, val, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, num_iter=bench.NUM_ITER)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    cases[name] = {'seconds_best_of_5': best, 'grid_points_per_s': bench.POINTS_PER_FIELD / best}
out['cases'] = cases
print(json.dumps(out, indent=1))
