import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
from fastbarnes import _lib
from oracle import oracle as orc
rng = np.random.default_rng(7)
for (n_outer, L, n_inner, T, n, alpha) in [(1, 1000, 64, 38, 4, 0.3), (1, 1000, 64, 38, 3, 0.3), (1, 1000, 64, 38, 2, 0.3),
                                           (1, 1000, 64, 38, 7, 0.3), (1, 1000, 16, 38, 4, 0.3), (1, 333, 64, 38, 4, 0.3),
                                           (1, 1000, 64, 8, 4, 0.3), (1, 1000, 64, 30, 4, 0.3), (1, 1000, 64, 38, 5, 0.3), (1, 1000, 64, 38, 6, 0.3)]:
    x = rng.normal(size=(n_outer, L, n_inner))
    y = x.copy()
    _lib.check(_lib.lib().fb_accumulate_lines_host(_lib.dptr(y), n_outer, L, n_inner, 2 * T + 1, n, alpha))
    bad = []
    for i in range(n_inner):
        line = np.ascontiguousarray(x[0, :, i])
        ref = orc._accumulate_tail_array(line.copy(), np.empty(L), L, 2 * T + 1, n, alpha)
        mism = np.nonzero(~((y[0, :, i] == ref) | (np.isnan(ref) & np.isnan(y[0, :, i]))))[0]
        if len(mism): bad.append((i, len(mism), int(mism[0]), int(mism[-1])))
    print((n_outer, L, n_inner, T, n), 'bad lines:', len(bad), bad[:4])
