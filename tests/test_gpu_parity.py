# -*- coding: utf-8 -*-
"""
GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(include/fastbarnes_b200.h) via the package's ctypes layer; results are compared with the
committed golden fixtures (outputs of the unmodified reference, oracle/gen_golden.py) and with
the CPU oracle on seeded inputs.

Tolerances: fp64 path -- BIT-EXACT (float64 quotient and float32 field), stronger than the
1e-12 relative of the north star.  S2 path: coordinates go through CUDA's tan/pow/sin/cos
(<= 2 ulp, not libm), so the float32 fields may differ from the reference by rounding flips:
tolerance 2e-6 relative (~16 float32 ulp at 1000 hPa is never reached; measured max is 1 ulp).
"""
import hashlib
from math import exp, sqrt, pi

import numpy as np
import pytest

from conftest import load_golden, bits_equal, same_up_to_zero_sign, CASES

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope='module')
def fb():
    from fastbarnes import interpolation
    return interpolation


@pytest.fixture(scope='module')
def fbS2():
    from fastbarnes import interpolationS2
    return interpolationS2


@pytest.fixture(scope='module')
def orc():
    from oracle import oracle
    return oracle


# ---------------------------------------------------------------------------------------------
# line kernel: the reference's own known-answer tests (tests/AccumulationTest.py:88-165)

def test_accumulate_tail_array_1_fold(fb):
    size = 32
    h_arr = np.empty(size)
    in_arr = np.zeros(size); in_arr[10] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 1, 0.25)
    assert np.all(out[:6] == 0) and np.all(out[15:] == 0)
    assert np.array_equal(out[6:15], [0.25, 1, 1, 1, 1, 1, 1, 1, 0.25])
    in_arr = np.zeros(size); in_arr[2] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 1, 0.25)
    assert np.array_equal(out[:7], [1, 1, 1, 1, 1, 1, 0.25]) and np.all(out[7:] == 0)
    in_arr = np.zeros(size); in_arr[30] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 1, 0.25)
    assert np.all(out[:26] == 0) and np.array_equal(out[26:], [0.25, 1, 1, 1, 1, 1])


def test_accumulate_tail_array_2_fold(fb):
    size = 32
    h_arr = np.empty(size)
    in_arr = np.zeros(size); in_arr[16] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 2, 0.5)
    assert np.all(out[:8] == 0) and np.all(out[25:] == 0)
    assert np.array_equal(out[8:25], [0.25, 1, 2, 3, 4, 5, 6, 7, 7.5, 7, 6, 5, 4, 3, 2, 1, 0.25])
    in_arr = np.zeros(size); in_arr[2] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 2, 0.5)
    assert np.array_equal(out[:11], [4.5, 5.5, 6.25, 6.5, 6, 5, 4, 3, 2, 1, 0.25]) and np.all(out[11:] == 0)
    in_arr = np.zeros(size); in_arr[30] = 1
    out = fb._accumulate_tail_array(in_arr, h_arr, size, 7, 2, 0.5)
    assert np.all(out[:22] == 0) and np.array_equal(out[22:], [0.25, 1, 2, 3, 4, 5, 5.5, 5.5, 5.25, 4.5])


def test_accumulate_array_folds(fb):
    size = 32
    h_arr = np.empty(size)
    in_arr = np.zeros(size); in_arr[10] = 1
    out = fb._accumulate_array(in_arr, h_arr, size, 9, 1)
    assert np.all(out[:6] == 0) and np.all(out[6:15] == 1) and np.all(out[15:] == 0)
    in_arr = np.zeros(size); in_arr[16] = 1
    out = fb._accumulate_array(in_arr, h_arr, size, 9, 2)
    assert np.array_equal(out[8:25], [1, 2, 3, 4, 5, 6, 7, 8, 9, 8, 7, 6, 5, 4, 3, 2, 1])
    in_arr = np.zeros(size); in_arr[2] = 1
    out = fb._accumulate_array(in_arr, h_arr, size, 9, 2)
    assert np.array_equal(out[:11], [5, 6, 7, 7, 7, 6, 5, 4, 3, 2, 1]) and np.all(out[11:] == 0)
    in_arr = np.zeros(size); in_arr[30] = 1
    out = fb._accumulate_array(in_arr, h_arr, size, 9, 2)
    assert np.all(out[:22] == 0) and np.array_equal(out[22:], [1, 2, 3, 4, 5, 6, 6, 6, 6, 5])


def test_accumulate_array_versions(fb):
    size = 32
    h_arr = np.empty(size)
    in_arr = np.zeros(size)
    in_arr[3] = 1; in_arr[9] = 2.5; in_arr[21] = -1.25; in_arr[30] = 1
    a = fb._accumulate_array(np.copy(in_arr), h_arr, size, 9, 3).copy()
    b = fb._accumulate_tail_array(np.copy(in_arr), h_arr, size, 9, 3, 0).copy()
    assert np.array_equal(a, b)
    c = fb._accumulate_tail_array(np.copy(in_arr), h_arr, size, 7, 3, 1).copy()
    assert np.array_equal(a, c)


def test_line_kernel_golden(fb):
    g = load_golden('kat_lines')
    i = 0
    while 'in_%d' % i in g:
        L, T, n, alpha = g['par_%d' % i]
        L, T, n = int(L), int(T), int(n)
        out = fb._accumulate_tail_array(g['in_%d' % i].copy(), np.empty(L), L, 2 * T + 1, n, alpha)
        assert same_up_to_zero_sign(out, g['tail_%d' % i]), 'tail line %d' % i
        out = fb._accumulate_array(g['in_%d' % i].copy(), np.empty(L), L, 2 * T + 1, n)
        assert same_up_to_zero_sign(out, g['plain_%d' % i]), 'plain line %d' % i
        i += 1
    assert i >= 9


def test_line_batches_vs_oracle(fb, orc):
    """ many lines at once, every fused-pass count, ragged inner sizes (lane masking) """
    from fastbarnes import _lib
    rng = np.random.default_rng(7)
    for (n_outer, L, n_inner, T, n, alpha) in [(3, 97, 37, 5, 4, 0.3), (1, 300, 16, 27, 4, 0.2083), (2, 64, 1, 0, 3, 0.03),
                                               (1, 130, 50, 13, 6, 0.7), (2, 61, 17, 27, 1, 0.5), (1, 260, 33, 54, 4, 0.926),
                                               (1, 90, 20, 4, 9, 0.11), (1, 1000, 5, 7, 2, 0.9), (1, 70, 3, 33, 5, 0.4)]:
        x = rng.normal(size=(n_outer, L, n_inner))
        x[rng.uniform(size=x.shape) < 0.3] = 0.0
        y = x.copy()
        _lib.check(_lib.lib().fb_accumulate_lines_host(_lib.dptr(y), n_outer, L, n_inner, 2 * T + 1, n, alpha))
        for o in range(n_outer):
            for i in range(n_inner):
                line = np.ascontiguousarray(x[o, :, i])
                ref = orc._accumulate_tail_array(line.copy(), np.empty(L), L, 2 * T + 1, n, alpha)
                assert same_up_to_zero_sign(y[o, :, i], ref), (n_outer, L, n_inner, T, n, o, i)


# ---------------------------------------------------------------------------------------------
# stages and whole path against the golden fixtures (outputs of the reference itself)

@pytest.mark.parametrize('name', CASES)
def test_inject_golden(fb, name):
    g = load_golden(name)
    size = tuple(int(s) for s in g['size'])
    vg, wg, offset = fb._inject_data(g['pts'], g['val'], g['x0'] * np.ones(len(size)), g['step'] * np.ones(len(size)), size)
    assert offset == float(g['offset'])
    assert bits_equal(vg, g['vin']) and bits_equal(wg, g['win'])


@pytest.mark.parametrize('name', CASES)
def test_convolve_golden(fb, name):
    g = load_golden(name)
    size = tuple(int(s) for s in g['size'])
    dim = len(size)
    vg, wg = g['vin'].copy(), g['win'].copy()
    sigma = g['sigma'] * np.ones(dim)
    step = g['step'] * np.ones(dim)
    ks = 2 * g['T'] + 1
    mdw = exp(-float(g['max_dist']) ** 2 / 2)
    if int(g['plain']):
        (fb._convolve_1d, fb._convolve_2d, fb._convolve_3d)[dim - 1](vg, wg, sigma, step, size, ks, int(g['num_iter']), mdw)
    else:
        (fb._convolve_tail_1d, fb._convolve_tail_2d, fb._convolve_tail_3d)[dim - 1](
            vg, wg, sigma, step, size, ks, int(g['num_iter']), g['alpha'], mdw)
    assert same_up_to_zero_sign(vg, g['vg'])
    assert same_up_to_zero_sign(wg, g['wg'])


@pytest.mark.parametrize('name', CASES)
def test_barnes_golden(fb, name):
    g = load_golden(name)
    size = tuple(int(s) for s in g['size'])
    dim = len(size)
    pts = g['pts'] if dim > 1 else g['pts'].reshape(-1)
    sigma = g['sigma'] if len(g['sigma']) > 1 else float(g['sigma'][0])
    x0 = g['x0'] if len(g['x0']) > 1 else float(g['x0'][0])
    step = g['step'] if len(g['step']) > 1 else float(g['step'][0])
    val0 = g['val'].copy()
    out32, out64 = fb.barnes(pts, g['val'], sigma, x0, step, size if dim > 1 else size[0],
                             method='convolution' if int(g['plain']) else 'optimized_convolution',
                             num_iter=int(g['num_iter']), max_dist=float(g['max_dist']), return_float64=True)
    assert np.array_equal(val0, g['val'])                     # caller's array untouched
    assert out32.dtype == np.float32 and out32.shape == size[::-1]
    assert bits_equal(out64, g['out64']), 'fp64 quotient must be bit-identical'
    assert bits_equal(out32, g['out32']), 'float32 field must be bit-identical'
    assert np.isnan(out32).mean() > 0.1                       # the mask is exercised


def test_paper_case_c1(fb):
    """ 2400x1200, N=3490, sigma=1, step=1/32, n=4: sha256 of the reference's output. """
    g = load_golden('c1_paper')
    size = tuple(int(s) for s in g['size'])
    out32, out64 = fb.barnes(g['pts'], g['val'], float(g['sigma']), g['x0'], float(g['step']), size,
                             num_iter=int(g['num_iter']), return_float64=True)
    assert out32.shape == (1200, 2400)
    assert bits_equal(out32[::13, ::17], g['sub_out32'])
    assert bits_equal(out64[::13, ::17], g['sub_out64'])
    assert sha(out32) == str(g['sha_out32'])
    assert sha(out64) == str(g['sha_out64'])
    assert np.isnan(out32).mean() == float(g['nan_frac'])
    vg, wg, offset = fb._inject_data(g['pts'], g['val'], g['x0'], float(g['step']) * np.ones(2), size)
    assert offset == float(g['offset'])
    assert sha(vg) == str(g['sha_vin']) and sha(wg) == str(g['sha_win'])


def test_random_cases_vs_oracle(fb, orc):
    """ seeded inputs at sizes the oracle finishes in seconds; 1D/2D/3D, both methods """
    rng = np.random.default_rng(99)
    for dim, size, sig in [(1, (3000,), [2.0]), (2, (333, 217), [1.3, 0.9]), (3, (70, 45, 50), [0.6, 0.5, 0.55]),
                           (2, (1000, 64), [6.0, 0.3]), (3, (33, 200, 17), [0.35, 2.0, 0.3])]:
        for n in (1, 2, 3, 4, 5, 6, 7, 10):
            for method in ('optimized_convolution', 'convolution'):
                N = 700
                ext = (np.asarray(size) - 1) * 0.1
                pts = rng.uniform(-0.05, 0.8, (N, dim)) * ext
                pts[:100] = pts[100:200]                      # duplicates -> ordered accumulation
                val = rng.normal(3, 20, N)
                try:
                    a = fb.barnes(pts if dim > 1 else pts[:, 0], val, sig, [0.0] * dim, 0.1, size if dim > 1 else size[0],
                                  method=method, num_iter=n)
                except RuntimeError as e:
                    assert 'kernel size' in str(e)
                    continue
                b = orc.barnes(pts, val, sig, [0.0] * dim, 0.1, size, method=method, num_iter=n, nthreads=8)
                assert bits_equal(a, b), (dim, size, n, method)


def test_batched_equals_singles(fb):
    rng = np.random.default_rng(5)
    size = (160, 120)
    counts = [300, 1, 57, 800, 300]
    offs = np.concatenate([[0], np.cumsum(counts)])
    pts = rng.uniform(0, 1, (offs[-1], 2)) * np.asarray([15.9, 11.9])
    val = rng.normal(1000, 10, offs[-1])
    out = fb.barnes_batched(pts, val, 0.8, [0.0, 0.0], 0.1, size, sample_offsets=offs, num_iter=4)
    assert out.shape == (5, 120, 160)
    for b in range(5):
        single = fb.barnes(pts[offs[b]:offs[b + 1]], val[offs[b]:offs[b + 1]], 0.8, [0.0, 0.0], 0.1, size, num_iter=4)
        assert bits_equal(out[b], single), b
    # (B, N, M) form
    p3 = rng.uniform(0, 1, (4, 200, 2)) * np.asarray([15.9, 11.9])
    v3 = rng.normal(0, 1, (4, 200))
    out = fb.barnes_batched(p3, v3, 0.8, [0.0, 0.0], 0.1, size)
    for b in range(4):
        assert bits_equal(out[b], fb.barnes(p3[b], v3[b], 0.8, [0.0, 0.0], 0.1, size))


def test_chunked_host_pipeline_and_kernel_variants(fb):
    """ tuning switches never change bits: multi-stream chunking of the host entry (ragged fields,
    partial last chunk) and q kernels vs first-generation sweep kernels """
    from fastbarnes import _lib
    L = _lib.lib()
    rng = np.random.default_rng(21)
    size = (200, 136)
    counts = [120, 400, 33, 250, 90, 310, 64]
    offs = np.concatenate([[0], np.cumsum(counts)])
    pts = rng.uniform(0, 1, (offs[-1], 2)) * np.asarray([19.9, 13.5])
    pts[:50] = pts[50:100]
    val = rng.normal(10, 3, offs[-1])
    ref = fb.barnes_batched(pts, val, 1.2, [0.0, 0.0], 0.1, size, sample_offsets=offs, num_iter=4)
    try:
        _lib.check(L.fb_set_option(b'host_chunk_fields', 2))
        a = fb.barnes_batched(pts, val, 1.2, [0.0, 0.0], 0.1, size, sample_offsets=offs, num_iter=4)
        p3 = pts[:offs[1]][None].repeat(6, axis=0) + rng.uniform(0, 0.01, (6, counts[0], 1))
        v3 = rng.normal(0, 1, (6, counts[0]))
        b = fb.barnes_batched(p3, v3, 1.2, [0.0, 0.0], 0.1, size, num_iter=5)
        _lib.check(L.fb_set_option(b'host_chunk_fields', 4))
        b_ref = fb.barnes_batched(p3, v3, 1.2, [0.0, 0.0], 0.1, size, num_iter=5)
        # 13 fields in chunks of 4: a first chunk of one field (a quarter), three full chunks
        p13 = pts[:offs[1]][None].repeat(13, axis=0) + rng.uniform(0, 0.01, (13, counts[0], 1))
        v13 = rng.normal(0, 1, (13, counts[0]))
        q4 = fb.barnes_batched(p13, v13, 1.2, [0.0, 0.0], 0.1, size, num_iter=4)
        _lib.check(L.fb_set_option(b'host_chunk_fields', 64))
        q_ref = fb.barnes_batched(p13, v13, 1.2, [0.0, 0.0], 0.1, size, num_iter=4)
        _lib.check(L.fb_set_option(b'host_chunk_fields', 4))
        _lib.check(L.fb_set_option(b'sweepq', 0)); _lib.check(L.fb_set_option(b'sweepp', 0))
        c = fb.barnes_batched(pts, val, 1.2, [0.0, 0.0], 0.1, size, sample_offsets=offs, num_iter=4)
    finally:
        L.fb_set_option(b'host_chunk_fields', 16)
        L.fb_set_option(b'sweepq', 1); L.fb_set_option(b'sweepp', 1)
    assert bits_equal(a, ref) and bits_equal(c, ref) and bits_equal(b, b_ref) and bits_equal(q4, q_ref)
    for i in range(len(counts)):
        single = fb.barnes(pts[offs[i]:offs[i + 1]], val[offs[i]:offs[i + 1]], 1.2, [0.0, 0.0], 0.1, size, num_iter=4)
        assert bits_equal(ref[i], single), i


# ---------------------------------------------------------------------------------------------
# the reference's integration tests (tests/BasicTest.py), through the drop-in API

def _in_range(arr, lo, hi, eps):
    return bool(np.all(np.logical_or(np.logical_and(arr >= lo - eps, arr <= hi + eps), np.isnan(arr))))


def test_basic_1d(fb):
    step, size = 1.0 / 64, 129
    pts = np.asarray([0.15, 0.2, 0.33, 0.35, 0.46, 0.59, 0.61, 0.66, 0.83, 0.98, 1.21, 1.29, 1.4, 1.57, 1.6, 1.79])
    val = np.asarray([1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 0.0, 1.0, 1.0, 1.0, 0.0, 1.0])
    for method in ['convolution', 'optimized_convolution']:
        for sigma in [0.5, 0.2, 0.05]:
            for num_iter in [1, 2, 3, 4, 5, 6, 8, 10]:
                res = fb.barnes(pts, val, sigma, 0.0, step, size, method=method, num_iter=num_iter)
                assert _in_range(res, 0.0, 1.0, 1e-15)


def test_basic_2d_3d(fb):
    rng = np.random.default_rng(3)
    step = 1.0 / 64
    pts = 2.0 * rng.random((60, 2))
    val = rng.integers(2, size=60).astype(np.float64)
    for method in ['convolution', 'optimized_convolution']:
        for sigma in [0.5, 0.2, 0.05]:
            for num_iter in [1, 2, 3, 4, 5, 6, 8, 10]:
                res = fb.barnes(pts, val, sigma, 0.0, step, (129, 129), method=method, num_iter=num_iter)
                assert _in_range(res, 0.0, 1.0, 5e-13)
    pts = 2.0 * rng.random((307, 3))
    val = rng.integers(2, size=307).astype(np.float64)
    for method in ['convolution', 'optimized_convolution']:
        for sigma in [0.5, 0.2, 0.05]:
            for num_iter in [1, 3, 4, 6, 10]:
                res = fb.barnes(pts, val, sigma, 0.0, step, (129, 129, 129), method=method, num_iter=num_iter)
                assert _in_range(res, 0.0, 1.0, 1e-12)


def test_array_width_vs_kernel_size(fb):
    sigma, step = 0.5, 1.0 / 32
    pts = np.asarray([0.2, 0.4, 0.7])
    val = np.asarray([18.5, 17.0, 19.25])
    for method, hk in (('convolution', fb.get_half_kernel_size), ('optimized_convolution', fb.get_half_kernel_size_opt)):
        for num_iter in [3, 4, 6]:
            kernel_size = 2 * hk(sigma, step, num_iter) + 1
            fb.barnes(pts, val, sigma, 0.0, step, kernel_size + 1, method=method, num_iter=num_iter)
            with pytest.raises(RuntimeError):
                fb.barnes(pts, val, sigma, 0.0, step, kernel_size, method=method, num_iter=num_iter)


def test_gaussian_1_dim_opt(fb):
    """ tests/GaussianApproximationTest.py:66-100 through _convolve_tail_1d """
    for sigma in [0.6, 0.8, 1.0, 1.5, 2.0]:
        for (num_iter, max_diff) in [[3, 0.02395], [4, 0.01405], [5, 0.012325], [6, 0.010045]]:
            delta = 1.0 / 512.0
            width = 3.5 * sigma
            grid = np.arange(-width, width + 0.000001, delta)
            gaussian = 1.0 / sqrt(2.0 * pi) / sigma * np.exp(-(grid / sigma) ** 2 / 2)
            values = np.zeros(len(grid))
            values[len(grid) // 2] = 1.0 / delta
            weights = np.copy(values)
            np_sigma, np_delta = fb._to_np(sigma), fb._to_np(delta)
            kernel_size = 2 * fb._get_half_kernel_size_opt(np_sigma, np_delta, num_iter) + 1
            tail_value = fb._get_tail_value(np_sigma, np_delta, num_iter)
            fb._convolve_tail_1d(values, weights, np_sigma, np_delta, (len(grid),), kernel_size, num_iter, tail_value, 9999.9)
            values *= (delta / 2 / sqrt(3 / num_iter) / sigma) ** num_iter
            assert np.max(np.abs(gaussian - values)) <= max_diff / sigma
            assert np.all(np.isnan(weights))


def test_gaussian_1_dim_plain(fb):
    """ tests/GaussianApproximationTest.py:32-64 through _convolve_1d """
    for sigma in [0.6, 1.0, 2.0]:
        for (num_iter, max_diff) in [[3, 0.02395], [4, 0.01405], [5, 0.01232], [6, 0.01004]]:
            delta = 1.0 / 512.0
            width = 3.5 * sigma
            grid = np.arange(-width, width + 0.000001, delta)
            sigma_eff = fb.get_sigma_effective(sigma, delta, num_iter)
            gaussian = 1.0 / sqrt(2.0 * pi) / sigma_eff * np.exp(-(grid / sigma_eff) ** 2 / 2)
            values = np.zeros(len(grid))
            values[len(grid) // 2] = 1.0 / delta
            weights = np.copy(values)
            np_sigma, np_delta = fb._to_np(sigma), fb._to_np(delta)
            kernel_size = 2 * fb._get_half_kernel_size(np_sigma, np_delta, num_iter) + 1
            fb._convolve_1d(values, weights, np_sigma, np_delta, (len(grid),), kernel_size, num_iter, 9999.9)
            values *= (delta / 2 / sqrt(3 / num_iter) / sigma_eff) ** num_iter
            assert np.max(np.abs(gaussian - values)) <= max_diff / sigma_eff


# ---------------------------------------------------------------------------------------------
# edge cases and size-independent properties

def test_edge_cases(fb):
    size = (64, 48)
    # sample exactly on the last node (xc == size-1) is skipped -> all NaN
    res = fb.barnes(np.asarray([[7.875, 2.0]]), np.asarray([5.0]), 0.5, [0.0, 0.0], 0.125, size)
    assert np.all(np.isnan(res))
    # a single sample gives a constant field wherever defined
    res = fb.barnes(np.asarray([[3.0, 2.0]]), np.asarray([5.0]), 0.5, [0.0, 0.0], 0.125, size)
    assert np.all(res[~np.isnan(res)] == 5.0) and (~np.isnan(res)).sum() > 100
    # out-of-grid samples are ignored
    a = fb.barnes(np.asarray([[3.0, 2.0], [-1.0, 2.0], [3.0, 99.0]]), np.asarray([5.0, 5.0, 5.0]), 0.5, [0.0, 0.0], 0.125, size)
    assert bits_equal(a, res)
    # float32 pts and integer val are converted
    b = fb.barnes(np.asarray([[3.0, 2.0]], dtype=np.float32), np.asarray([5]), 0.5, [0.0, 0.0], 0.125, size)
    assert bits_equal(b, res)
    # N == 0
    with pytest.raises(ValueError):
        fb.barnes(np.zeros((0, 2)), np.zeros(0), 0.5, [0.0, 0.0], 0.1, size)
    # validation errors are RuntimeErrors like in the reference
    with pytest.raises(RuntimeError):
        fb.barnes([[1.0, 2.0]], np.asarray([1.0]), 0.5, [0.0, 0.0], 0.1, size)
    with pytest.raises(RuntimeError):
        fb.barnes(np.zeros((2, 2)), np.zeros(3), 0.5, [0.0, 0.0], 0.1, size)
    with pytest.raises(RuntimeError):
        fb.barnes(np.zeros((2, 2)), np.zeros(2), [0.5, 0.5, 0.5], [0.0, 0.0], 0.1, size)
    with pytest.raises(RuntimeError):
        fb.barnes(np.zeros((2, 2)), np.zeros(2), 0.5, [0.0, 0.0], 0.1, size, method='bogus')


def test_properties_full_size(fb):
    """ size-independent properties at the paper grid (2400 x 1200) with 50k random samples """
    rng = np.random.default_rng(11)
    size = (2400, 1200)
    step = 1.0 / 32
    x0 = np.asarray([-26.0 + step, 34.5])
    N = 50000
    pts = x0 + rng.uniform(0, 1, (N, 2)) * np.asarray([(2400 - 1) / 32, (1200 - 1) / 32])
    val = rng.normal(1000, 10, N)
    a = fb.barnes(pts, val, 1.0, x0, step, size)
    # determinism: bit-identical on repetition
    assert bits_equal(a, fb.barnes(pts, val, 1.0, x0, step, size))
    # constant observations reproduce the constant exactly
    c = fb.barnes(pts, np.full(N, 1013.25), 1.0, x0, step, size)
    assert np.all(c[~np.isnan(c)] == np.float32(1013.25))
    assert np.array_equal(np.isnan(c), np.isnan(a))
    # convex combination: min <= field <= max
    assert np.nanmin(a) >= val.min() - 1e-3 and np.nanmax(a) <= val.max() + 1e-3
    # affine equivariance with exactly representable scale/shift of the centred values
    b = fb.barnes(pts, 2.0 * val, 1.0, x0, step, size)
    assert np.allclose(b[~np.isnan(b)], 2.0 * a[~np.isnan(a)], rtol=3e-7, atol=0)


def test_1d_exact_default_and_segmented_option(fb, orc):
    """ 1D: the default walks the single line with the 2 x num_iter (field, pass) chains side by side
    (csrc/fb_line1d.cuh) and is BIT-IDENTICAL to the reference at any length, for every pass count, kernel width
    (T = 1 .. 55) and for lines whose length is no multiple of the chunk.  exact=False opts into overlapping segments
    swept in parallel: exact in exact arithmetic; vs the reference the fp64 quotient then differs by the reference's
    own accumulated rounding (tolerance 1e-8 * value range at 2^18 points, NaN mask identical, float32 within 8 ulp). """
    rng = np.random.default_rng(1234)
    L = 2 ** 18
    N = L // 64
    pts = rng.uniform(0, L - 1, N)
    val = rng.normal(0, 1, N)
    vrange = val.max() - val.min()
    for n, sigma in ((4, 32.0), (6, 32.0), (5, 3.0), (1, 10.0), (2, 64.0), (3, 1.2)):
        ref = orc._interpolate_opt_convol(pts.reshape(-1, 1), val.copy(), np.asarray([sigma]), np.zeros(1), np.ones(1),
                                          (L,), n, exp(-3.5 ** 2 / 2), stages=True)
        e32, e64 = fb.barnes(pts, val, sigma, 0.0, 1.0, L, num_iter=n, return_float64=True)          # default: exact
        assert bits_equal(e64, ref['out64']) and bits_equal(e32, ref['out32']), (n, sigma)
        if n in (4, 1):
            a32, a64 = fb.barnes(pts, val, sigma, 0.0, 1.0, L, num_iter=n, return_float64=True, exact=False)
            assert np.array_equal(np.isnan(a64), np.isnan(ref['out64']))
            m = ~np.isnan(ref['out64'])
            assert np.max(np.abs(a64[m] - ref['out64'][m])) <= 1e-8 * vrange, (n, sigma)
            assert np.mean(a32[m] != ref['out32'][m]) < 1e-3
            ulp = np.abs(a32[m].view(np.int32).astype(np.int64) - ref['out32'][m].view(np.int32).astype(np.int64))
            assert ulp.max() <= 8
    # odd lengths, the old lane-pair walk as a cross-check, short grids
    from fastbarnes import _lib
    for Ls in (5000, 1025, 77777):
        p2 = rng.uniform(0, Ls - 1, Ls // 25)
        v2 = rng.normal(5, 2, Ls // 25)
        a = fb.barnes(p2, v2, 12.0, 0.0, 1.0, Ls)
        assert bits_equal(a, orc.barnes(p2.reshape(-1, 1), v2, 12.0, 0.0, 1.0, (Ls,))), Ls
        try:
            _lib.check(_lib.lib().fb_set_option(b'line1d', 0))
            assert bits_equal(a, fb.barnes(p2, v2, 12.0, 0.0, 1.0, Ls)), Ls
        finally:
            _lib.lib().fb_set_option(b'line1d', 1)
    seg = fb.barnes(p2, v2, 12.0, 0.0, 1.0, Ls, exact=False)               # forcing segments works too
    ex = fb.barnes(p2, v2, 12.0, 0.0, 1.0, Ls)
    assert np.array_equal(np.isnan(seg), np.isnan(ex)) and np.nanmax(np.abs(seg - ex)) <= 1e-5


def test_1d_exact_2e22_points_is_fast(fb, orc):
    """ BASELINE configs[1] at 1/16 of its length: N = 65536 samples on a 2^22-point line, sigma 32 (T = 27), n = 4:
    bit-identical to the oracle, and the device pipeline stays below 60 ms (the lane-pair walk needed ~590 ms, the
    reference's Numba code ~290 ms for this length) """
    import time
    rng = np.random.default_rng(1234)
    L = 2 ** 22
    N = L // 64
    pts = rng.uniform(0, L - 1, N)
    val = rng.normal(0, 1, N)
    ref = orc.barnes(pts.reshape(-1, 1), val, 32.0, 0.0, 1.0, (L,), num_iter=4, nthreads=8)
    a = fb.barnes(pts, val, 32.0, 0.0, 1.0, L, num_iter=4)
    assert bits_equal(a, ref)
    t0 = time.perf_counter()
    fb.barnes(pts, val, 32.0, 0.0, 1.0, L, num_iter=4)
    dt = time.perf_counter() - t0
    assert dt < 0.060, dt


@pytest.mark.parametrize('sig', [[0.9, 0.8, 0.7], [1.3, 1.2, 1.1]], ids=['planes_T2', 'interleaved_T3'])
def test_z_slab_decomposition_single_gpu(fb, sig):
    """ 3D z-slabs, emulated on one GPU: own planes are injected and x/y-swept per slab (boundary planes
    first, as the multi-GPU run does to overlap the exchange), halo planes copied between the slabs, fused z sweep per slab.
    Tolerance: the x/y stages are bit-identical; the z sweep restarts its accumulator at the halo edge, so the fp64
    quotient may differ at rounding level: |difference| <= 1e-12 * (range of the values) -- stated against the value
    range, not against the quotient itself, because a relative bound is meaningless where the field crosses zero
    (zero-centred values are one of the cases) -- NaN mask identical, float32 equal on > 99.9 % of the points and never
    off by more than 1 ulp where |field| > 1e-3 * range.  Slabs thinner than the halo (16 slabs of 7-8 planes, halo 12)
    take their halo planes from several ranks.  The two kernel widths run the two forms of the slab buffers: planes of
    values and of weights (T=2: first-generation kernels) and interleaved nodes (T=3: q kernels). """
    from fastbarnes import distributed
    rng = np.random.default_rng(8)
    size = (96, 80, 120)
    N = 4000
    pts = rng.uniform(0.0, 1.0, (N, 3)) * (np.asarray(size) - 1) * 0.25
    pts[:200] = pts[200:400]
    probe = distributed.BarnesSlab3D(sig, [0.0, 0.0, 0.0], 0.25, size, N, num_iter=4, nslabs=1, slab=0)
    assert probe.interleaved == (sig[0] > 1.0)
    del probe
    for centre in (280.0, 0.0):
        val = rng.normal(centre, 7.0, N)
        vrange = float(val.max() - val.min())
        ref32, ref64 = fb.barnes(pts, val, sig, [0.0, 0.0, 0.0], 0.25, size, num_iter=4, return_float64=True)
        for nslabs in (1, 2, 4, 16):
            got32, got64 = distributed.barnes_slabs_emulated(pts, val, sig, [0.0, 0.0, 0.0], 0.25, size, nslabs, num_iter=4,
                                                             want_float64=True)
            assert got32.shape == ref32.shape == (120, 80, 96)
            assert np.array_equal(np.isnan(got32), np.isnan(ref32))
            m = ~np.isnan(ref64)
            err = np.max(np.abs(got64[m] - ref64[m]))
            assert err <= 1e-12 * vrange, (centre, nslabs, err)
            if nslabs == 1:
                assert bits_equal(got64, ref64) and bits_equal(got32, ref32)
            assert np.mean(got32[m] != ref32[m]) < 1e-3
            big = m & (np.abs(ref64) > 1e-3 * vrange)
            assert np.max(np.abs(got32[big].view(np.int32).astype(np.int64) - ref32[big].view(np.int32).astype(np.int64))) <= 1


@pytest.mark.parametrize('sig', [[0.9, 0.8, 0.7], [1.3, 1.2, 1.1]], ids=['planes_T2', 'interleaved_T3'])
def test_z_slab_planes_before_the_z_sweep_are_bit_identical(fb, sig):
    """ What a slab holds after injection and the x / y sweeps -- per-plane work -- must equal the same planes of the
    undivided run bit for bit: the slab injects from the samples compacted to its neighbourhood (in order: nodes with
    several records, here 200 repeated locations and 300 samples in one cell, are summed in sample order), and only the z
    sweep, which restarts its accumulator at the halo edge, may differ by rounding. """
    torch = pytest.importorskip('torch')
    from fastbarnes import distributed
    rng = np.random.default_rng(88)
    size = (96, 80, 120)
    N = 6000
    pts = rng.uniform(-0.02, 1.02, (N, 3)) * (np.asarray(size) - 1) * 0.25
    pts[:200] = pts[200:400]
    pts[400:700] = (np.asarray([40.2, 33.6, 57.4]) + rng.uniform(0, 0.6, (300, 3))) * 0.25
    val = rng.normal(3.0, 7.0, N)
    dp, dv = torch.from_numpy(pts).cuda(), torch.from_numpy(val).cuda()

    def planes(nslabs, r):
        s = distributed.BarnesSlab3D(sig, [0.0, 0.0, 0.0], 0.25, size, N, num_iter=4, nslabs=nslabs, slab=r)
        s.inject(dp, dv)
        s.sweeps(0, s.zc)
        v, w = s.own_planes()
        return s.z0, s.z1, v.contiguous().cpu().numpy(), w.contiguous().cpu().numpy()

    _, _, ref_v, ref_w = planes(1, 0)
    for nslabs in (3, 16):
        for r in range(nslabs):
            z0, z1, v, w = planes(nslabs, r)
            assert np.array_equal(v.view(np.uint64), ref_v[z0:z1].view(np.uint64)), (nslabs, r)
            assert np.array_equal(w.view(np.uint64), ref_w[z0:z1].view(np.uint64)), (nslabs, r)


# ---------------------------------------------------------------------------------------------
# S2 path

S2_RTOL = 2e-6


def test_lambert_to_map(fbS2):
    from fastbarnes.util import lambert_conformal
    g = load_golden('s2_res8')
    proj = fbS2.get_lambert_proj()
    assert np.array_equal(np.asarray(proj), g['proj'])        # host libm: bit-identical
    lam = lambert_conformal.to_map(g['pts'], g['pts'].copy(), *proj)
    assert np.max(np.abs(lam - g['lam_pts'])) <= 1e-13        # CUDA math: <= 2 ulp per function


@pytest.mark.parametrize('res', [8, 32])
def test_barnes_s2_golden(fbS2, res):
    g = load_golden('s2_res%d' % res)
    size = tuple(int(s) for s in g['size'])
    step = float(g['step'])
    out = fbS2.barnes_S2(g['pts'], g['val'], 1.0, g['x0'], step, size, method='optimized_convolution_S2', num_iter=4)
    lam = fbS2.barnes_S2(g['pts'], g['val'], 1.0, g['x0'], step, size, method='optimized_convolution_S2', num_iter=4,
                         resample=False)
    assert out.shape == (size[1], size[0]) and out.dtype == np.float32
    assert lam.shape == (int(44.0 / step), int(64.0 / step))
    ref_out = g['out'] if 'out' in g else None
    if ref_out is not None:
        assert np.array_equal(np.isnan(out), np.isnan(ref_out))
        m = ~np.isnan(ref_out)
        assert np.max(np.abs(out[m] - ref_out[m]) / np.abs(ref_out[m])) <= S2_RTOL
        assert np.mean(out[m] != ref_out[m]) < 0.01
        assert np.array_equal(np.isnan(lam), np.isnan(g['lam']))
        ml = ~np.isnan(g['lam'])
        assert np.max(np.abs(lam[ml] - g['lam'][ml]) / np.abs(g['lam'][ml])) <= S2_RTOL
    so, sl = out[::7, ::11], lam[::7, ::11]
    assert np.array_equal(np.isnan(so), np.isnan(g['sub_out']))
    m = ~np.isnan(g['sub_out'])
    assert np.max(np.abs(so[m] - g['sub_out'][m]) / np.abs(g['sub_out'][m])) <= S2_RTOL
    ml = ~np.isnan(g['sub_lam'])
    assert np.max(np.abs(sl[ml] - g['sub_lam'][ml]) / np.abs(g['sub_lam'][ml])) <= S2_RTOL
    # split API (timing5 of the reference) gives the same field as the fused call
    p1 = fbS2.interpolate_opt_convol_S2_part1(g['pts'], g['val'].copy(), np.full(2, 1.0), g['x0'], np.full(2, step), size,
                                              4, exp(-3.5 ** 2 / 2))
    out2 = fbS2.interpolate_opt_convol_S2_part2(*p1)
    assert bits_equal(out2, out) and bits_equal(p1[0], lam)


def test_s2_res64_vs_oracle(fbS2, orc):
    """ BASELINE configs[3]: 4800x2400 lon/lat grid, Lambert grid 4096x2816, T=54 (wide rings) """
    g = load_golden('c1_paper')
    step = 1.0 / 64
    x0 = np.asarray([-26.0 + step, 34.5])
    size = (4800, 2400)
    out = fbS2.barnes_S2(g['pts'], g['val'], 1.0, x0, step, size, method='optimized_convolution_S2', num_iter=4)
    ref = orc.barnes_S2(g['pts'], g['val'], 1.0, x0, step, size, num_iter=4, nthreads=8)
    assert out.shape == ref.shape == (2400, 4800)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.max(np.abs(out[m] - ref[m]) / np.abs(ref[m])) <= S2_RTOL
    assert np.mean(out[m] != ref[m]) < 0.01


def test_s2_user_map(fbS2, orc):
    """ generalised S2: a user-chosen Lambert map (North America) against the fixture composed
    from the reference's own functions, and the automatic map against the oracle run on the same map """
    g = load_golden('s2_map_na')
    size = tuple(int(s) for s in g['size'])
    n = int(g['num_iter'])
    lmap = fbS2.LambertMap(g['proj'], g['lam_x0'], g['lam_extent'])
    out = fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, method='optimized_convolution_S2',
                         num_iter=n, lambert_map=lmap)
    lam = fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, method='optimized_convolution_S2',
                         num_iter=n, resample=False, lambert_map=lmap)
    for a, b in ((out, g['out']), (lam, g['lam'])):
        assert a.shape == b.shape and a.dtype == np.float32
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = ~np.isnan(b)
        assert np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) <= S2_RTOL
        assert np.mean(a[m] != b[m]) < 0.01
    # the default map is the reference's
    d = fbS2.LambertMap.default()
    assert d.proj == tuple(fbS2.get_lambert_proj()) and d.lam_x0 == (-32.0, -2.0) and d.lam_extent == (64.0, 44.0)
    # automatic map: window covers the target grid (no NaN from leaving the window where the oracle has data)
    auto = fbS2.LambertMap.for_grid(g['x0'], g['step'], size, margin=3.5)
    out_a = fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, method='optimized_convolution_S2',
                           num_iter=n, lambert_map=auto)
    ref_a = orc.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, num_iter=n, nthreads=4,
                          lambert_map=(auto.proj, auto.lam_x0, auto.lam_extent))
    assert np.array_equal(np.isnan(out_a), np.isnan(ref_a))
    m = ~np.isnan(ref_a)
    assert m.mean() > 0.95
    assert np.max(np.abs(out_a[m] - ref_a[m]) / np.abs(ref_a[m])) <= S2_RTOL
    out_s = fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, method='optimized_convolution_S2',
                           num_iter=n, lambert_map='auto')
    assert bits_equal(out_s, out_a)
    # two maps of the same region give nearly the same field (conformal maps, sigma small against the domain)
    both = m & ~np.isnan(out)
    assert np.sqrt(np.mean((out_a[both] - out[both]) ** 2)) < 0.1
    # a window that does not cover the grid gives NaN outside instead of reading out of bounds
    small = fbS2.LambertMap(g['proj'], (-10.0, -5.0), (20.0, 10.0))
    out_w = fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, method='optimized_convolution_S2',
                           num_iter=n, lambert_map=small)
    assert 0.5 < np.isnan(out_w).mean() < 1.0
    with pytest.raises(RuntimeError):
        fbS2.barnes_S2(g['pts'], g['val'], g['sigma'], g['x0'], g['step'], size, lambert_map='nope')


def test_nan_and_constant_values(fb):
    """ np.amin/np.amax propagate NaN: one NaN observation makes the whole field NaN (reference
    behaviour of _normalize_values); identical observations give offset == value and 0/w + value """
    rng = np.random.default_rng(4)
    pts = rng.uniform(1, 5, (50, 2))
    val = rng.normal(0, 1, 50)
    val[7] = np.nan
    res = fb.barnes(pts, val, 0.5, [0.0, 0.0], 0.1, (64, 64))
    assert np.all(np.isnan(res))
    res = fb.barnes(pts, np.full(50, -3.25), 0.5, [0.0, 0.0], 0.1, (64, 64))
    assert np.all(res[~np.isnan(res)] == np.float32(-3.25)) and (~np.isnan(res)).sum() > 500


def test_resample_exact_on_reference_field(fbS2):
    """ resampling alone, fed with the reference's own Lambert field: the only non-libm inputs are
    the per-row rho and per-column sin/cos tables """
    g = load_golden('s2_res8')
    size = tuple(int(s) for s in g['size'])
    out = fbS2._resample(g['lam'], np.asarray([-32.0, -2.0]), g['x0'], np.full(2, float(g['step'])), size, *g['proj'])
    m = ~np.isnan(g['out'])
    assert np.array_equal(np.isnan(out), np.isnan(g['out']))
    assert np.max(np.abs(out[m] - g['out'][m]) / np.abs(g['out'][m])) <= S2_RTOL


# ---------------------------------------------------------------------------------------------
# exact Gaussian sums ("next" row N3): 'naive' / 'radius' / 'naive_S2'

from test_oracle_golden import EXACT_CASES, exact_case, close_exact      # noqa: E402


@pytest.mark.parametrize('name,method,s2', EXACT_CASES)
def test_exact_methods_golden(fb, fbS2, orc, name, method, s2):
    """ reference interpolation.py:862-938 (naive), :809-855 (radius), interpolationS2.py:260-301
    (naive_S2) through the public API: against the reference's own output (fixture) and against
    the oracle, to rounding (sums in sample order, CUDA libm vs glibc; tolerance in close_exact:
    |a-b| <= 1e-11 + 1e-12 |b|, NaN masks identical). """
    g = load_golden('exact_methods')
    pts, val, sigma, x0, step, size, kw = exact_case(g, name)
    val_before = val.copy()
    out = (fbS2.barnes_S2 if s2 else fb.barnes)(pts, val, sigma, x0, step, size, method=method, **kw)
    assert np.array_equal(val, val_before)
    assert out.dtype == np.float64 and out.shape == size[::-1]
    assert close_exact(out, g[name + '_out'])
    ref = (orc.barnes_S2 if s2 else orc.barnes)(pts, val, sigma, x0, step, size, method=method, nthreads=4, **kw)
    assert close_exact(out, ref)


def test_exact_methods_agree_with_convolution(fb):
    """ the yardstick use: the optimized convolution with n=4 approximates the exact sum; their RMSE on the
    paper's samples at a coarse grid is small compared with the spread of the values, 'radius' equals
    'naive' where the truncated weights are negligible, and a many-sample tile loop is covered """
    g = load_golden('exact_methods')
    pts, val = g['paper_naive_pts'], g['paper_naive_val']
    x0, step, size = [-25.5, 34.5], 0.5, (150, 75)
    naive = fb.barnes(pts, val, 1.0, x0, step, size, method='naive')
    radius = fb.barnes(pts, val, 1.0, x0, step, size, method='radius')
    conv = fb.barnes(pts, val, 1.0, x0, step, size, method='optimized_convolution', num_iter=6)
    m = ~np.isnan(radius) & ~np.isnan(conv)
    assert m.mean() > 0.8
    assert np.median(np.abs(radius[m] - naive[m])) < 0.01         # min_weight = 0.001 truncation
    rmse = np.sqrt(np.mean((conv[m] - naive[m]) ** 2))
    assert rmse < 0.5 and rmse < 0.05 * np.std(val)


# ---------------------------------------------------------------------------------------------
# fp32 working precision (north_star's fp32 path): tolerance-level parity, stated here

def fp32_close(a, b, val_spread):
    """
    Tolerance of the fp32 path against the fp64 path (which is bit-identical to the reference):
      * NaN masks differ on at most 1e-5 of the points (points whose weight is within fp32 rounding of the
        max_dist threshold),
      * RMS difference <= 1e-5 * spread of the values (a few float32 ulps of a field of magnitude 1000),
      * max difference <= 1e-3 * spread (reached only next to the max_dist boundary, where the weights are 3e-4
        of those inside clusters of samples).
    """
    assert a.shape == b.shape and a.dtype == b.dtype == np.float32
    mism = np.isnan(a) != np.isnan(b)
    assert mism.mean() <= 1e-5, mism.sum()
    m = ~np.isnan(a) & ~np.isnan(b)
    d = np.abs(a[m].astype(np.float64) - b[m].astype(np.float64))
    assert np.sqrt(np.mean(d ** 2)) <= 1e-5 * val_spread, np.sqrt(np.mean(d ** 2))
    assert d.max() <= 1e-3 * val_spread, d.max()
    return True


def test_fp32_path_paper_case(fb):
    g = load_golden('c1_paper')
    step = 1.0 / 32
    x0 = np.asarray([-26.0 + step, 34.5])
    size = (2400, 1200)
    spread = float(g['val'].max() - g['val'].min())
    naive = None
    for method in ('optimized_convolution', 'convolution'):
        for n in (2, 4, 6):
            a = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, method=method, num_iter=n, precision='fp32')
            b = fb.barnes(g['pts'], g['val'], 1.0, x0, step, size, method=method, num_iter=n)
            assert fp32_close(a, b, spread)
    # "matching RMSE against the naive method": the fp32 field is as close to the exact Gaussian sum as the fp64 one
    naive = fb.barnes(g['pts'], g['val'], 1.0, x0, 0.25, (300, 150), method='naive')
    a = fb.barnes(g['pts'], g['val'], 1.0, x0, 0.25, (300, 150), num_iter=4, precision='fp32')        # T = 3
    b = fb.barnes(g['pts'], g['val'], 1.0, x0, 0.25, (300, 150), num_iter=4)
    m = ~np.isnan(a) & ~np.isnan(b)
    rmse32 = np.sqrt(np.mean((a[m] - naive[m]) ** 2))
    rmse64 = np.sqrt(np.mean((b[m] - naive[m]) ** 2))
    assert abs(rmse32 - rmse64) <= 1e-3 * rmse64


def test_fp32_path_shapes_batches_3d(fb):
    rng = np.random.default_rng(77)
    pts = rng.uniform(0, 1, (20000, 2)) * [70, 35]
    val = rng.normal(1000, 10, 20000)
    spread = float(val.max() - val.min())
    for size, sigma, step in (((1000, 517), 1.0, 1 / 16), ((333, 1201), 0.6, 1 / 16), ((70, 33), 1.5, 0.25), ((2001, 97), 0.8, 1 / 16)):
        a = fb.barnes(pts, val, sigma, [0.0, 0.0], step, size, precision='fp32')
        b = fb.barnes(pts, val, sigma, [0.0, 0.0], step, size)
        assert fp32_close(a, b, spread)
    # anisotropic sigma / step
    a = fb.barnes(pts, val, [1.2, 0.7], [0.0, 0.0], [0.125, 0.0625], (500, 500), num_iter=3, precision='fp32')
    b = fb.barnes(pts, val, [1.2, 0.7], [0.0, 0.0], [0.125, 0.0625], (500, 500), num_iter=3)
    assert fp32_close(a, b, spread)
    # batched == singles, bit for bit, also in fp32
    B, N = 5, 3000
    bp = rng.uniform(0, 1, (B, N, 2)) * [30, 20]
    bv = rng.normal(0, 1, (B, N))
    out = fb.barnes_batched(bp, bv, 1.0, [0.0, 0.0], 0.125, (241, 161), precision='fp32')
    for i in range(B):
        assert bits_equal(out[i], fb.barnes(bp[i], bv[i], 1.0, [0.0, 0.0], 0.125, (241, 161), precision='fp32'))
    # 3D
    p3 = rng.uniform(0, 1, (50000, 3)) * [60, 50, 40]
    v3 = rng.normal(0, 1, 50000)
    a = fb.barnes(p3, v3, 3.0, [0.0, 0.0, 0.0], 0.5, (121, 101, 81), precision='fp32')
    b = fb.barnes(p3, v3, 3.0, [0.0, 0.0, 0.0], 0.5, (121, 101, 81))
    assert fp32_close(a, b, float(v3.max() - v3.min()))
    # what the fp32 path does not cover is refused, not approximated
    with pytest.raises(RuntimeError, match='2D and 3D'):
        fb.barnes(rng.uniform(0, 10, 100), rng.normal(0, 1, 100), 1.0, 0.0, 0.1, 101, precision='fp32')
    with pytest.raises(RuntimeError, match='fp32 path does not cover'):
        fb.barnes(pts, val, 0.1, [0.0, 0.0], 1 / 16, (200, 200), precision='fp32')       # T = 1
    with pytest.raises(RuntimeError, match="'fp64' or 'fp32'"):
        fb.barnes(pts, val, 1.0, [0.0, 0.0], 1 / 16, (200, 200), precision='fp16')


# ---------------------------------------------------------------------------------------------
# large batches (device-resident, like bench.py's): the q kernels against the oracle and against the first-generation
# shared-memory kernel

@pytest.mark.parametrize('ratio,iters', [(8, (2, 3, 4, 5, 6)), (32, (4,))])
def test_large_batch_kernels(fb, orc, ratio, iters):
    """ device-resident batches (the host entry point works in chunks of 4 fields and never gets there) """
    torch = pytest.importorskip('torch')
    from fastbarnes import _lib
    L = _lib.lib()
    rng = np.random.default_rng(1000 + ratio)
    F, N = 40, 1500
    size = (512, 500)                                   # 32 line groups per field and sweep: 1280 work items
    step = 1.0 / ratio
    ext = np.asarray([(size[0] - 1) * step, (size[1] - 1) * step])
    pts = rng.uniform(0, 1, (F, N, 2)) * ext
    pts[:, :100] = pts[:, 100:200]                      # repeated locations
    val = rng.normal(1000, 10, (F, N))
    d_pts = torch.from_numpy(pts.reshape(F * N, 2)).cuda()
    d_val = torch.from_numpy(val.reshape(F * N)).cuda()

    def run(method, n, precision='fp64'):
        plan = fb.BarnesDevice(2, 1.0, [0.0, 0.0], step, size, nfields=F, nsamples=F * N, method=method, num_iter=n,
                               precision=precision)
        n0 = L.fb_kernel_launch_count()
        out = plan(d_pts, d_val).cpu().numpy()
        return out, L.fb_kernel_launch_count() - n0

    for n in iters:
        for method in (('optimized_convolution', 'convolution') if n == 4 else ('optimized_convolution',)):
            out, _ = run(method, n)
            try:
                _lib.check(L.fb_set_option(b'sweepq', 0)); _lib.check(L.fb_set_option(b'sweepp', 0))
                smem, _ = run(method, n)
            finally:
                L.fb_set_option(b'sweepq', 1); L.fb_set_option(b'sweepp', 1)
            assert bits_equal(out, smem), (n, method)
            for i in (0, 17, F - 1):
                ref = orc.barnes(pts[i], val[i], 1.0, [0.0, 0.0], step, size, method=method, num_iter=n, nthreads=4)
                assert bits_equal(out[i], ref), (n, method, i)
    # the fp32 path on the same batch
    a, _ = run('optimized_convolution', 4, 'fp32')
    b, _ = run('optimized_convolution', 4)
    assert fp32_close(a, b, float(val.max() - val.min()))


def test_bench_shape_against_oracle(fb, orc):
    """ The exact workload of bench.py (BASELINE configs[1]): 2400x1200 grid at 1/32 degree, 50000 samples
    per field, sigma 1 degree, 4 passes, the samples bench.make_fields draws.  One field through the host API and one
    64-field sub-batch device-resident (what one timed launch of the bench processes): fields 0, 31 and 63 bit for bit
    against the oracle. """
    torch = pytest.importorskip('torch')
    import bench
    pts, val = bench.make_fields(0, bench.SUB_FIELDS)
    refs = {i: orc.barnes(pts[i], val[i], bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, num_iter=bench.NUM_ITER, nthreads=8)
            for i in (0, 31, bench.SUB_FIELDS - 1)}
    one = fb.barnes(pts[0], val[0], bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, num_iter=bench.NUM_ITER)
    assert one.shape == (1200, 2400) and bits_equal(one, refs[0])
    F, N = bench.SUB_FIELDS, bench.N_PER_FIELD
    plan = fb.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=F, nsamples=F * N, num_iter=bench.NUM_ITER)
    out = plan(torch.from_numpy(pts.reshape(F * N, 2)).cuda(), torch.from_numpy(val.reshape(F * N)).cuda())
    for i, ref in refs.items():
        assert bits_equal(out[i].cpu().numpy(), ref), i


def test_injection_lists_vs_segments(fb, orc):
    """ Interleaved fp64 nodes (the form the q sweeps read): the two-pass injection that links the
    records of a node into a list against the three-pass count / allocate / place version and the oracle,
    bit for bit -- random samples, repeated locations (lists of 25 records: heap sort; of 2: insertion sort)
    and fields with every sample in a single cell or at a single location (lists of 1500 records). """
    torch = pytest.importorskip('torch')
    from fastbarnes import _lib
    L = _lib.lib()
    rng = np.random.default_rng(77)
    F, N = 40, 1500
    size = (512, 500)
    step = 0.125
    ext = np.asarray([(size[0] - 1) * step, (size[1] - 1) * step])
    pts = rng.uniform(0, 1, (F, N, 2)) * ext
    pts[0, :100] = pts[0, 100:200]                          # pairs
    pts[1, :600] = pts[1, 600:625].repeat(24, axis=0)       # 25 records on each node of 25 cells
    pts[2] = (np.asarray([200.25, 300.75]) + rng.uniform(0, 0.5, (N, 2))) * step   # one cell
    pts[3, :] = pts[3, 0]                                   # one location
    val = rng.normal(1000, 10, (F, N))
    d_pts = torch.from_numpy(pts.reshape(F * N, 2)).cuda()
    d_val = torch.from_numpy(val.reshape(F * N)).cuda()
    plan = fb.BarnesDevice(2, 1.0, [0.0, 0.0], step, size, nfields=F, nsamples=F * N, num_iter=4, want_float64=True)
    out = plan(d_pts, d_val).cpu().numpy()
    out64 = plan.out64.cpu().numpy()
    try:
        _lib.check(L.fb_set_option(b'inject_lists', 0))
        seg = plan(d_pts, d_val).cpu().numpy()
        seg64 = plan.out64.cpu().numpy()
        _lib.check(L.fb_set_option(b'sweepq', 0)); _lib.check(L.fb_set_option(b'sweepp', 0))         # planes of values / weights, first-generation sweeps
        planes = plan(d_pts, d_val).cpu().numpy()
    finally:
        L.fb_set_option(b'inject_lists', 1)
        L.fb_set_option(b'sweepq', 1); L.fb_set_option(b'sweepp', 1)
    assert bits_equal(out, seg) and bits_equal(out, planes)
    assert np.array_equal(out64.view(np.uint64), seg64.view(np.uint64))
    for i in (0, 1, 2, 3, F - 1):
        ref = orc.barnes(pts[i], val[i], 1.0, [0.0, 0.0], step, size, num_iter=4, nthreads=4)
        assert bits_equal(out[i], ref), i
    # ragged fields (one empty, one with a single sample) and samples outside the grid
    counts = rng.integers(800, 1500, F)
    counts[4] = 0
    counts[5] = 1
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rp = rng.uniform(-0.05, 1.05, (int(offs[-1]), 2)) * ext          # ~10 % of the samples are skipped
    rv = rng.normal(5.0, 2.0, int(offs[-1]))
    plan = fb.BarnesDevice(2, 1.0, [0.0, 0.0], step, size, nfields=F, nsamples=int(offs[-1]), num_iter=4,
                           sample_offsets=offs)
    out = plan(torch.from_numpy(rp).cuda(), torch.from_numpy(rv).cuda()).cpu().numpy()
    assert np.isnan(out[4]).all()
    for i in (0, 5, 6, F - 1):
        a, b = int(offs[i]), int(offs[i + 1])
        ref = orc.barnes(rp[a:b], rv[a:b], 1.0, [0.0, 0.0], step, size, num_iter=4, nthreads=4)
        assert bits_equal(out[i], ref), i


def test_3d_volume_kernels(fb, orc):
    """ a volume with >= 1184 line groups in every sweep: all three kernel modes (transposing x sweep, in-place y sweep,
    finalising z sweep) against the oracle, bit for bit """
    rng = np.random.default_rng(31)
    size = (160, 128, 160)
    step = 0.25
    pts = rng.uniform(0, 1, (60000, 3)) * (np.asarray(size) - 1) * step
    pts[:500] = pts[500:1000]
    val = rng.normal(5.0, 2.0, 60000)
    for n in (4, 5):
        out = fb.barnes(pts, val, 1.0, [0.0, 0.0, 0.0], step, size, num_iter=n)
        ref = orc.barnes(pts, val, 1.0, [0.0, 0.0, 0.0], step, size, num_iter=n, nthreads=8)
        assert bits_equal(out, ref), n
    a = fb.barnes(pts, val, 1.0, [0.0, 0.0, 0.0], step, size, num_iter=4, precision='fp32')
    assert fp32_close(a, fb.barnes(pts, val, 1.0, [0.0, 0.0, 0.0], step, size, num_iter=4), float(val.max() - val.min()))


# ---------------------------------------------------------------------------------------------
# second-generation sweep kernels (csrc/fb_sweepq.cuh): one warp = 16 lines x all passes, TMA row staging, rings in tensor
# and shared memory.  They are the default for 2D / 3D fp64 grids whose kernels have 2T+2 >= 8 elements on every axis.

def _q_option(value, small_batch=0):
    """ sweepq on / off; the small-batch kernels (fb_sweepp.cuh), which would take over for the few fields of these
    tests, are off unless asked for (1: when the batch is small -- the default of the library, 2: always) """
    from fastbarnes import _lib
    _lib.check(_lib.lib().fb_set_option(b'sweepq', value))
    _lib.check(_lib.lib().fb_set_option(b'sweepp', small_batch))


@pytest.mark.parametrize('dim,size,ratio,nf,iters', [
    (2, (512, 500), (8.0, 8.0), 4, (1, 2, 3, 4, 5, 6)),     # T = 5..13: 2T+2 no multiple of 8 (ring mirrors), rings in tensor memory
    (2, (333, 217), (13.0, 9.0), 3, (3, 4)),                # ragged 16-line groups on both axes, anisotropic kernel
    (2, (1000, 300), (60.0, 5.0), 2, (4,)),                 # T = 51 (four warps per CTA) and T = 3 (2T+2 = 8, the minimum)
    (2, (640, 400), (40.0, 34.0), 2, (2, 6)),               # T = 48 / 27: two and six passes
    (3, (160, 128, 96), (6.0, 7.0, 6.5), 1, (1, 2, 4, 5)),  # 3D: transposing x sweep, in-place y sweep, finalising z sweep
])
def test_sweepq_vs_first_generation(fb, dim, size, ratio, nf, iters):
    """ bit for bit against the first-generation kernels (themselves bit-identical to the reference): float32 field and
    fp64 quotient, samples outside the grid, repeated locations """
    torch = pytest.importorskip('torch')
    rng = np.random.default_rng(4242 + dim + len(iters))
    step = 0.1
    sigma = [r * step for r in ratio]
    ext = (np.asarray(size) - 1) * step
    N = 4000
    pts = rng.uniform(-0.02, 1.02, (nf, N, dim)) * ext
    pts[:, :200] = pts[:, 200:400]
    val = rng.normal(100, 20, (nf, N))
    d_pts = torch.from_numpy(pts.reshape(nf * N, dim)).cuda()
    d_val = torch.from_numpy(val.reshape(nf * N)).cuda()
    for n in iters:
        plan = fb.BarnesDevice(dim, sigma, [0.0] * dim, step, size, nfields=nf, nsamples=nf * N, num_iter=n, want_float64=True)
        try:
            _q_option(1)
            a, a64 = plan(d_pts, d_val).cpu().numpy(), plan.out64.cpu().numpy()
            _q_option(0)
            b, b64 = plan(d_pts, d_val).cpu().numpy(), plan.out64.cpu().numpy()
            _q_option(1, small_batch=2)
            c, c64 = plan(d_pts, d_val).cpu().numpy(), plan.out64.cpu().numpy()
        finally:
            _q_option(1, small_batch=1)
        assert bits_equal(a, b) and bits_equal(a64, b64), (dim, size, n)
        assert bits_equal(a, c) and bits_equal(a64, c64), ('small-batch kernels', dim, size, n)


@pytest.mark.parametrize('T', [28, 40, 54, 59])
def test_sweepq_wide_kernels_vs_oracle(fb, orc, T):
    """ kernels wider than the 27 the first-generation tensor-memory kernel stopped at (S2 at resolution 64 has T = 54):
    the q kernels keep them on chip with four warps per CTA; against the oracle, bit for bit """
    n = 4
    s = sqrt(((2 * T + 1.5) ** 2 - 1) * n / 12.0)          # sigma / step that gives this half kernel size
    assert fb.get_half_kernel_size_opt(s, 1.0, n) == T
    rng = np.random.default_rng(T)
    size = (2 * T + 40, 2 * T + 61)
    pts = rng.uniform(0, 1, (3000, 2)) * (np.asarray(size) - 1)
    pts[:100] = pts[100:200]
    val = rng.normal(1000, 10, 3000)
    ref = orc.barnes(pts, val, s, [0.0, 0.0], 1.0, size, num_iter=n, nthreads=8)
    try:
        for small_batch in (0, 2):                          # the q kernels, then the small-batch kernels (the default for one field)
            _q_option(1, small_batch)
            a = fb.barnes(pts, val, s, [0.0, 0.0], 1.0, size, num_iter=n)
            assert bits_equal(a, ref), small_batch
    finally:
        _q_option(1, small_batch=1)


@pytest.mark.parametrize('size,ratio,n,nf,N', [
    ((88, 83, 58), (7.3, 7.1, 4.6), 2, 5, 2096),        # T = (7, 7, 4): in-place y sweep of a small volume, five fields
    ((72, 53, 77), (2.9, 7.6, 13.2), 1, 5, 239),        # one pass: the y sweep writes into the injection buffer
])
def test_small_volumes_many_times(fb, orc, size, ratio, n, nf, N):
    """ Regression (found by tools/fuzz_parity.py): the transposing / in-place q sweeps request their next staging chunk as
    soon as the current one is in registers.  A first version of that request was not ordered behind ALL the loads; the
    TMA unit then overwrote rows still being read, and small 3D volumes differed from the reference in about a third of
    the runs (never the 2D bench batch).  Twenty runs of two such volumes, every field bit for bit. """
    rng = np.random.default_rng(670)
    step = 0.25
    sigma = [r * step for r in ratio]
    pts = rng.uniform(-0.03, 1.03, (nf, N, 3)) * (np.asarray(size) - 1) * step
    pts[:, :60] = pts[:, 60:120]
    val = rng.normal(100.0, 20.0, (nf, N))
    refs = [orc.barnes(pts[i], val[i], sigma, [0.0] * 3, step, size, num_iter=n, nthreads=8) for i in range(nf)]
    try:
        _q_option(1, small_batch=0)
        for run in range(20):
            out = fb.barnes_batched(pts, val, sigma, [0.0] * 3, step, size, num_iter=n)
            for i in range(nf):
                assert bits_equal(out[i], refs[i]), (run, i)
    finally:
        _q_option(1, small_batch=1)


def test_random_configurations_against_oracle(fb, orc):
    """ A fixed-seed slice of tools/fuzz_parity.py: random 2D / 3D grids, kernel widths, pass counts, sample and field counts
    through the default kernel choice, the q kernels alone, the pass-parallel kernels and the first-generation kernel. """
    rng = np.random.default_rng(20261018)
    done = 0
    try:
        while done < 40:
            dim = int(rng.choice([2, 2, 3]))
            n = int(rng.integers(1, 7))
            size = tuple(int(x) for x in (rng.integers(40, 300, 2) if dim == 2 else rng.integers(24, 80, 3)))
            step = float(rng.choice([0.1, 0.25, 1.0]))
            ratio = rng.uniform(1.2, 12.0, dim) if dim == 3 else rng.uniform(1.2, 30.0, dim)
            sigma = [float(r * step) for r in ratio]
            if any(2 * fb.get_half_kernel_size_opt(sigma[m], step, n) + 1 >= size[m] for m in range(dim)):
                continue
            nf = int(rng.choice([1, 2, 5]))
            N = int(rng.integers(30, 2000))
            pts = rng.uniform(-0.03, 1.03, (nf, N, dim)) * (np.asarray(size) - 1) * step
            k = min(N // 3, 80)
            pts[:, :k] = pts[:, k:2 * k]
            val = rng.normal(rng.uniform(-50, 500), rng.uniform(0.1, 30), (nf, N))
            refs = [orc.barnes(pts[i], val[i], sigma, [0.0] * dim, step, size, num_iter=n, nthreads=8) for i in range(nf)]
            for q, sb in ((1, 1), (1, 0), (1, 2), (0, 0)):
                _q_option(q, small_batch=sb)
                out = fb.barnes_batched(pts, val, sigma, [0.0] * dim, step, size, num_iter=n)
                for i in range(nf):
                    assert bits_equal(out[i], refs[i]), (done, dim, size, n, nf, N, (q, sb), i)
            done += 1
    finally:
        _q_option(1, small_batch=1)


def test_spare_buffers_with_very_wide_kernels(fb, orc):
    """ Kernels so wide that their rings do not fit on chip: the grid takes the first-generation path, where a launch
    covers one pass (2D) or the y sweep of a volume is split into several launches that are NOT in place -- the later
    launches must write into the free buffer pair (advisor finding, round 1). """
    torch = pytest.importorskip('torch')
    rng = np.random.default_rng(222)
    # 2D: T_y = 449 -> one pass per launch on y (ping-pong between the buffer pairs)
    size, nf, N = (48, 912), 24, 600
    step = 1.0
    sig = [8.0, 520.0]
    pts = rng.uniform(0, 1, (nf, N, 2)) * (np.asarray(size) - 1)
    val = rng.normal(5, 2, (nf, N))
    plan = fb.BarnesDevice(2, sig, [0.0, 0.0], step, size, nfields=nf, nsamples=nf * N, num_iter=4)
    out = plan(torch.from_numpy(pts.reshape(nf * N, 2)).cuda(), torch.from_numpy(val.reshape(nf * N)).cuda()).cpu().numpy()
    for i in (0, nf - 1):
        ref = orc.barnes(pts[i], val[i], sig, [0.0, 0.0], step, size, num_iter=4, nthreads=8)
        assert bits_equal(out[i], ref), i
    # 3D: T_y = 222 with three passes -> the y sweep runs as 2 + 1 passes, the second launch is not in place
    size3 = (40, 460, 44)
    sig3 = [6.0, 223.0, 5.0]
    n3 = 3
    assert fb.get_half_kernel_size_opt(sig3[1], 1.0, n3) >= 222
    pts3 = rng.uniform(0, 1, (2000, 3)) * (np.asarray(size3) - 1)
    val3 = rng.normal(5, 2, 2000)
    a = fb.barnes(pts3, val3, sig3, [0.0, 0.0, 0.0], step, size3, num_iter=n3)
    ref = orc.barnes(pts3, val3, sig3, [0.0, 0.0, 0.0], step, size3, num_iter=n3, nthreads=8)
    assert bits_equal(a, ref)


def test_sparse_injection_feeds_the_x_sweep(fb, orc):
    """ opt-in path `sparse_inject`: the samples are binned by the x sweep's unit of work (sort by cell: count, scan,
    fill; one ordered segmented reduce per bucket) and producer warps turn the node entries into the rows the passes
    read -- no dense injection grid.  Bit for bit against the dense path and the oracle: repeated locations, a field with
    all samples in one cell (one bucket with 6000 records), samples outside the grid, ragged fields. """
    torch = pytest.importorskip('torch')
    from fastbarnes import _lib
    L = _lib.lib()
    rng = np.random.default_rng(909)
    F, N = 40, 1500
    size = (512, 500)
    step = 0.125
    ext = np.asarray([(size[0] - 1) * step, (size[1] - 1) * step])
    pts = rng.uniform(-0.03, 1.03, (F, N, 2)) * ext
    pts[0, :100] = pts[0, 100:200]
    pts[1, :600] = pts[1, 600:625].repeat(24, axis=0)
    pts[2] = (np.asarray([200.25, 300.75]) + rng.uniform(0, 0.5, (N, 2))) * step
    pts[3, :] = pts[3, 0]
    val = rng.normal(1000, 10, (F, N))
    counts = rng.integers(900, N + 1, F)
    counts[4] = 0
    counts[5] = 1
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rp = np.concatenate([pts[i, :counts[i]] for i in range(F)])
    rv = np.concatenate([val[i, :counts[i]] for i in range(F)])
    d_p, d_v = torch.from_numpy(rp).cuda(), torch.from_numpy(rv).cuda()
    for n in (4, 3):
        plan = fb.BarnesDevice(2, 1.0, [0.0, 0.0], step, size, nfields=F, nsamples=int(offs[-1]), num_iter=n, sample_offsets=offs,
                               want_float64=True)
        try:
            _lib.check(L.fb_set_option(b'sparse_inject', 1))
            a, a64 = plan(d_p, d_v).cpu().numpy(), plan.out64.cpu().numpy()
        finally:
            L.fb_set_option(b'sparse_inject', 0)
        b, b64 = plan(d_p, d_v).cpu().numpy(), plan.out64.cpu().numpy()
        assert bits_equal(a, b) and bits_equal(a64, b64), n
        assert np.isnan(a[4]).all()
        for i in (0, 1, 2, 3, 5, F - 1):
            lo, hi = int(offs[i]), int(offs[i + 1])
            ref = orc.barnes(rp[lo:hi], rv[lo:hi], 1.0, [0.0, 0.0], step, size, num_iter=n, nthreads=4)
            assert bits_equal(a[i], ref), (n, i)
