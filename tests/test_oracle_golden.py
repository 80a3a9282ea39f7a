# -*- coding: utf-8 -*-
"""
CPU tests (no GPU): the oracle (oracle/fb_oracle.c, a C restatement of the reference) against
  * the reference's own known-answer vectors (tests/AccumulationTest.py of the reference),
  * the committed golden fixtures produced by the unmodified reference (oracle/gen_golden.py),
  * the reference itself, imported live, when /root/reference exists (build container only).
All comparisons are bit-exact, except the exact-sum methods (naive / radius / naive_S2), which are
compared to rounding with the tolerance stated in close_exact().
"""
import hashlib
import os
import sys
from math import exp

import numpy as np
import pytest

from conftest import load_golden, bits_equal, CASES
from oracle import oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_reference_kat_tail_1_fold():
    # reference tests/AccumulationTest.py:88-113
    size = 32
    h = np.empty(size)
    x = np.zeros(size); x[10] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 1, 0.25)
    assert np.all(out[:6] == 0) and np.array_equal(out[6:15], [0.25, 1, 1, 1, 1, 1, 1, 1, 0.25]) and np.all(out[15:] == 0)
    x = np.zeros(size); x[2] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 1, 0.25)
    assert np.array_equal(out[:7], [1, 1, 1, 1, 1, 1, 0.25]) and np.all(out[7:] == 0)
    x = np.zeros(size); x[30] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 1, 0.25)
    assert np.all(out[:26] == 0) and np.array_equal(out[26:], [0.25, 1, 1, 1, 1, 1])


def test_reference_kat_tail_2_fold():
    # reference tests/AccumulationTest.py:115-140
    size = 32
    h = np.empty(size)
    x = np.zeros(size); x[16] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 2, 0.5)
    assert np.array_equal(out[8:25], [0.25, 1, 2, 3, 4, 5, 6, 7, 7.5, 7, 6, 5, 4, 3, 2, 1, 0.25])
    assert np.all(out[:8] == 0) and np.all(out[25:] == 0)
    x = np.zeros(size); x[2] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 2, 0.5)
    assert np.array_equal(out[:11], [4.5, 5.5, 6.25, 6.5, 6, 5, 4, 3, 2, 1, 0.25]) and np.all(out[11:] == 0)
    x = np.zeros(size); x[30] = 1
    out = orc._accumulate_tail_array(x, h, size, 7, 2, 0.5)
    assert np.all(out[:22] == 0) and np.array_equal(out[22:], [0.25, 1, 2, 3, 4, 5, 5.5, 5.5, 5.25, 4.5])


def test_reference_kat_plain_and_versions():
    # reference tests/AccumulationTest.py:34-86, :142-165
    size = 32
    h = np.empty(size)
    x = np.zeros(size); x[16] = 1
    out = orc._accumulate_array(x, h, size, 9, 2)
    assert np.array_equal(out[8:25], [1, 2, 3, 4, 5, 6, 7, 8, 9, 8, 7, 6, 5, 4, 3, 2, 1])
    x = np.zeros(size); x[2] = 1
    out = orc._accumulate_array(x, h, size, 9, 2)
    assert np.array_equal(out[:11], [5, 6, 7, 7, 7, 6, 5, 4, 3, 2, 1])
    x = np.zeros(size); x[3] = 1; x[9] = 2.5; x[21] = -1.25; x[30] = 1
    a = orc._accumulate_array(x.copy(), h, size, 9, 3).copy()
    b = orc._accumulate_tail_array(x.copy(), h, size, 9, 3, 0.0).copy()
    c = orc._accumulate_tail_array(x.copy(), h, size, 7, 3, 1.0).copy()
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_golden_lines():
    g = load_golden('kat_lines')
    i = 0
    while 'in_%d' % i in g:
        L, T, n, alpha = g['par_%d' % i]
        L, T, n = int(L), int(T), int(n)
        out = orc._accumulate_tail_array(g['in_%d' % i].copy(), np.empty(L), L, 2 * T + 1, n, alpha)
        assert bits_equal(out, g['tail_%d' % i])
        out = orc._accumulate_array(g['in_%d' % i].copy(), np.empty(L), L, 2 * T + 1, n)
        assert bits_equal(out, g['plain_%d' % i])
        i += 1
    assert i == 9


def test_golden_params():
    tab = load_golden('params')['table']
    mdw = exp(-3.5 ** 2 / 2)
    for sigma, step, n, T, Tp, alpha, csf, csfp in tab:
        n = int(n)
        assert orc._get_half_kernel_size_opt(sigma, step, n)[0] == int(T)
        assert orc._get_half_kernel_size(sigma, step, n)[0] == int(Tp)
        assert orc._get_tail_value(sigma, step, n)[0] == alpha
        assert orc.conv_scale_factor([2 * int(T) + 1], [alpha], [sigma], [step], n, mdw) == csf
        assert orc.conv_scale_factor([2 * int(Tp) + 1], [0.0], [sigma], [step], n, mdw) == csfp
    # the values SURVEY.md section 8c quotes for sigma=1, step=1/32
    want = {1: 54, 2: 38, 3: 31, 4: 27, 5: 24, 6: 22, 10: 17, 20: 11, 50: 7}
    for n, T in want.items():
        assert orc._get_half_kernel_size_opt(1.0, 1 / 32, n)[0] == T
    assert orc._get_tail_value(1.0, 1 / 32, 4)[0] == 0.20833333333333334


@pytest.mark.parametrize('name', CASES)
def test_golden_cases(name):
    g = load_golden(name)
    size = tuple(int(s) for s in g['size'])
    dim = len(size)
    st = orc._interpolate_opt_convol(g['pts'], g['val'].copy(), g['sigma'] * np.ones(dim), g['x0'] * np.ones(dim),
                                     g['step'] * np.ones(dim), size, int(g['num_iter']),
                                     exp(-float(g['max_dist']) ** 2 / 2), plain=bool(int(g['plain'])), nthreads=2,
                                     stages=True)
    assert st['offset'] == float(g['offset'])
    for key in ('vin', 'win', 'vg', 'wg', 'out64', 'out32'):
        assert bits_equal(st[key], g[key]), key


def test_golden_c1_paper():
    g = load_golden('c1_paper')
    size = tuple(int(s) for s in g['size'])
    st = orc._interpolate_opt_convol(g['pts'], g['val'].copy(), np.full(2, float(g['sigma'])), g['x0'],
                                     np.full(2, float(g['step'])), size, int(g['num_iter']), exp(-3.5 ** 2 / 2),
                                     nthreads=4, stages=True)
    assert st['offset'] == float(g['offset']) == 1007.6500000000001
    assert sha(st['out32']) == str(g['sha_out32']) and sha(st['out64']) == str(g['sha_out64'])
    assert sha(st['vin']) == str(g['sha_vin']) and sha(st['wg']) == str(g['sha_wg'])
    assert float(np.isnan(st['out32']).mean()) == float(g['nan_frac']) == 0.10696284722222223
    assert float(g['csf']) == 30240548.729689308 and list(g['T']) == [27, 27]


def test_golden_s2():
    g = load_golden('s2_res8')
    size = tuple(int(s) for s in g['size'])
    assert np.array_equal(np.asarray(orc.get_lambert_proj()), g['proj'])
    assert bits_equal(orc.to_map(g['pts'], g['pts'].copy(), *g['proj']), g['lam_pts'])
    out = orc.barnes_S2(g['pts'], g['val'], 1.0, g['x0'], float(g['step']), size, num_iter=4, nthreads=2)
    lam = orc.barnes_S2(g['pts'], g['val'], 1.0, g['x0'], float(g['step']), size, num_iter=4, resample=False, nthreads=2)
    assert bits_equal(out, g['out']) and bits_equal(lam, g['lam'])


def test_golden_s2_user_map():
    # generalised S2: the reference's own to_map / _interpolate_opt_convol / _resample composed
    # on a North-America Lambert map (oracle/gen_golden.py s2_map_case) vs the oracle with the same map
    g = load_golden('s2_map_na')
    size = tuple(int(s) for s in g['size'])
    lmap = (g['proj'], g['lam_x0'], g['lam_extent'])
    assert bits_equal(np.asarray(orc.create_proj(-97.625, 37.375, 29.125, 45.625)), g['proj'])
    out = orc.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, num_iter=int(g['num_iter']),
                        nthreads=2, lambert_map=lmap)
    lam = orc.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, num_iter=int(g['num_iter']),
                        resample=False, nthreads=2, lambert_map=lmap)
    assert bits_equal(lam, g['lam']) and bits_equal(out, g['out'])


EXACT_CASES = [  # name, method, S2
    ('naive_2d', 'naive', False), ('naive_2d_aniso', 'naive', False), ('radius_2d', 'radius', False),
    ('radius_2d_sparse', 'radius', False), ('radius_2d_minw', 'radius', False), ('naive_1d', 'naive', False),
    ('naive_3d', 'naive', False), ('naive_S2', 'naive_S2', True), ('paper_naive', 'naive', False),
    ('paper_radius', 'radius', False), ('paper_naive_S2', 'naive_S2', True)]


def exact_case(g, name):
    """ inputs of one exact_methods.npz case: (pts, val, sigma, x0, step, size, kwargs). """
    src = name if name + '_pts' in g else 'paper_naive'
    pts, val, args = g[src + '_pts'], g[src + '_val'], g[name + '_args']
    dim = pts.shape[1]
    sigma, x0, step = args[:dim], args[dim:2 * dim], args[2 * dim:3 * dim]
    size = tuple(int(v) for v in args[3 * dim:4 * dim])
    kw = dict(max_dist=2.0, min_weight=0.01) if name == 'radius_2d_minw' else {}
    return pts, val, sigma, x0, step, size, kw


def close_exact(a, b, atol=1e-11, rtol=1e-12):
    """ The exact-sum methods are compared to rounding: the reference sums with np.dot / np.sum (BLAS /
    pairwise order) or in kd-tree order, the restatements in sample order.  NaN masks must be identical. """
    assert a.shape == b.shape and a.dtype == b.dtype == np.float64
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan)
    return np.all(np.abs(a[~nan] - b[~nan]) <= atol + rtol * np.abs(b[~nan]))


@pytest.mark.parametrize('name,method,s2', EXACT_CASES)
def test_golden_exact_methods(name, method, s2):
    # 'naive' / 'radius' / 'naive_S2' of the reference (interpolation.py:862-938, :809-855,
    # interpolationS2.py:260-301) vs the oracle's sample-order sums
    g = load_golden('exact_methods')
    pts, val, sigma, x0, step, size, kw = exact_case(g, name)
    fn = orc.barnes_S2 if s2 else orc.barnes
    out = fn(pts, val, sigma, x0, step, size, method=method, nthreads=4, **kw)
    assert close_exact(out, g[name + '_out'])


def test_thread_count_does_not_change_bits():
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 10, (300, 2))
    val = rng.normal(0, 1, 300)
    a = orc.barnes(pts, val, 0.7, [0.0, 0.0], 0.1, (110, 105), nthreads=1)
    b = orc.barnes(pts, val, 0.7, [0.0, 0.0], 0.1, (110, 105), nthreads=4)
    assert bits_equal(a, b)


# ---------------------------------------------------------------------------------------------
REF = '/root/reference'


_LIVE_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import fastbarnes.interpolation as ref
rng = np.random.default_rng(17)
out = {}
i = 0
for dim, size in [(1, (400,)), (2, (150, 90)), (3, (40, 36, 30))]:
    for n, method in [(1, 'optimized_convolution'), (4, 'optimized_convolution'), (3, 'convolution')]:
        N = 250
        pts = rng.uniform(-0.05, 0.7, (N, dim)) * (np.asarray(size) - 1) * 0.1
        pts[:40] = pts[40:80]
        val = rng.normal(2, 5, N)
        sig = [0.8, 0.6, 0.5][:dim]
        r = ref.barnes(pts if dim > 1 else pts[:, 0], val, sig, [0.0] * dim, 0.1, size if dim > 1 else size[0],
                       method=method, num_iter=n)
        out['pts%d' % i] = pts; out['val%d' % i] = val; out['res%d' % i] = r
        out['par%d' % i] = np.asarray([dim, n, method == 'convolution'] + list(size))
        i += 1
np.savez(sys.argv[2], **out)
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'fastbarnes')), reason='reference checkout not present')
def test_oracle_vs_live_reference(tmp_path):
    """ differential check against the reference run under Numba in a subprocess (build container
    only; the reference package has the same import name as the drop-in, hence the subprocess) """
    import subprocess
    script = tmp_path / 'live_ref.py'
    script.write_text(_LIVE_SCRIPT)
    outfile = tmp_path / 'live.npz'
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    subprocess.check_call([sys.executable, str(script), REF, str(outfile)], env=env, cwd=str(tmp_path))
    z = np.load(outfile)
    i = 0
    while 'res%d' % i in z.files:
        par = z['par%d' % i]
        dim, n, plain = int(par[0]), int(par[1]), bool(par[2])
        size = tuple(int(v) for v in par[3:])
        sig = [0.8, 0.6, 0.5][:dim]
        o = orc.barnes(z['pts%d' % i], z['val%d' % i], sig, [0.0] * dim, 0.1, size,
                       method='convolution' if plain else 'optimized_convolution', num_iter=n, nthreads=2)
        assert bits_equal(z['res%d' % i], o), (dim, n, plain)
        i += 1
    assert i == 9
