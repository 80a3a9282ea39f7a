# -*- coding: utf-8 -*-
"""
gen_golden.py -- generates the golden fixtures under tests/golden/ by running the
UNMODIFIED reference (MeteoSwiss/fast-barnes-py, imported from /root/reference
under Numba).  Test infrastructure; runs only in the build container (the GPU box
has no /root/reference).  Re-run with:  python oracle/gen_golden.py

Fixtures (all inputs seeded, outputs produced by the reference's own functions):
  kat_lines.npz     random lines through _accumulate_tail_array / _accumulate_array
  params.npz        T / alpha / conv_scale_factor tables
  case_{1,2,3}d_*.npz  inputs + every stage of _interpolate_opt_convol
  c1_paper.npz      the paper case (2400x1200, N=3490): inputs, sha256 of the float32
                    output, a strided sub-sample of the fp64 quotient and float32 output
  s2_*.npz          barnes_S2 at step 1/8 (full arrays) and 1/32 (digest + sub-sample)
  exact_methods.npz 'naive', 'radius' and 'naive_S2' (the exact Gaussian sums) on small seeded
                    cases and on the paper's samples at a coarse grid
                    (only this one:  python oracle/gen_golden.py exact)
  s2_map_na.npz     the S2 pipeline composed from the reference's own functions (to_map,
                    _interpolate_opt_convol, _resample) on a user-chosen Lambert map over North
                    America instead of the hard-coded European one
                    (only this one:  python oracle/gen_golden.py s2map)
"""
import hashlib
import os
import sys
from math import exp, sqrt, pi

import numpy as np

REF = os.environ.get('FB_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
from numba import njit                                       # noqa: E402
from fastbarnes import interpolation as ref                  # noqa: E402
from fastbarnes import interpolationS2 as refS2              # noqa: E402
from fastbarnes.util import lambert_conformal as reflc       # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@njit
def _csf_ref(kernel_size, tail_value, sigma, step, num_iter, max_dist_weight):
    # the expression of fastbarnes/interpolation.py:424-425, compiled by Numba like the original
    conv_scale_factor = (kernel_size + 2 * tail_value) ** num_iter / sqrt(2 * pi) / (sigma / step)
    return np.prod(conv_scale_factor) * max_dist_weight


def ref_stages(pts, val, sigma, x0, step, size, num_iter, max_dist, plain=False):
    """ Replays _interpolate_opt_convol (interpolation.py:329-367) stage by stage. """
    dim = len(size)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, dim)
    val = val.copy()
    sigma = np.full(dim, sigma, dtype=np.float64) if np.isscalar(sigma) else np.asarray(sigma, np.float64)
    step = np.full(dim, step, dtype=np.float64) if np.isscalar(step) else np.asarray(step, np.float64)
    x0 = np.full(dim, x0, dtype=np.float64) if np.isscalar(x0) else np.asarray(x0, np.float64)
    mdw = exp(-max_dist ** 2 / 2)
    offset = ref._normalize_values(val)
    rsize = size[::-1]
    vg = np.zeros(rsize)
    wg = np.zeros(rsize)
    (ref._inject_data_1d, ref._inject_data_2d, ref._inject_data_3d)[dim - 1](vg, wg, pts, val, x0, step, size)
    vin, win = vg.copy(), wg.copy()
    if plain:
        T = ref._get_half_kernel_size(sigma, step, num_iter)
        ks = 2 * T + 1
        tv = np.zeros(dim)
        (ref._convolve_1d, ref._convolve_2d, ref._convolve_3d)[dim - 1](vg, wg, sigma, step, size, ks,
                                                                         num_iter, mdw)
    else:
        T = ref._get_half_kernel_size_opt(sigma, step, num_iter)
        ks = 2 * T + 1
        tv = ref._get_tail_value(sigma, step, num_iter)
        (ref._convolve_tail_1d, ref._convolve_tail_2d, ref._convolve_tail_3d)[dim - 1](
            vg, wg, sigma, step, size, ks, num_iter, tv, mdw)
    with np.errstate(all='ignore'):
        out64 = vg / wg + offset
    out32 = out64.astype(np.float32)
    return dict(offset=offset, vin=vin, win=win, vg=vg, wg=wg, out64=out64, out32=out32,
                T=T.astype(np.int32), alpha=tv, csf=_csf_ref(ks, tv, sigma, step, num_iter, mdw))


def main():
    rng = np.random.default_rng(20221017)

    # ---- line kernel -----------------------------------------------------------------
    lines = {}
    cases = [(32, 3, 1, 0.25), (32, 3, 2, 0.5), (64, 0, 4, 0.0333), (200, 27, 4, 0.2083333333333333),
             (131, 64, 3, 0.9), (58, 27, 6, 0.1), (1000, 13, 5, 0.61), (57, 27, 1, 0.3), (300, 54, 4, 0.926)]
    for i, (L, T, n, alpha) in enumerate(cases):
        x = rng.normal(size=L) * rng.uniform(0.1, 100)
        x[rng.uniform(size=L) < 0.5] = 0.0
        y = ref._accumulate_tail_array(x.copy(), np.empty(L), L, 2 * T + 1, n, alpha).copy()
        yp = ref._accumulate_array(x.copy(), np.empty(L), L, 2 * T + 1, n).copy()
        lines['in_%d' % i] = x
        lines['tail_%d' % i] = y
        lines['plain_%d' % i] = yp
        lines['par_%d' % i] = np.asarray([L, T, n, alpha], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, 'kat_lines.npz'), **lines)

    # ---- parameter tables --------------------------------------------------------------
    rows = []
    for sigma, step in [(1.0, 1 / 32), (1.0, 1 / 64), (0.5, 1 / 64), (0.2, 1 / 64), (0.05, 1 / 64), (0.5, 1.0),
                        (8.0, 1.0), (32.0, 1.0), (16.0, 1.0), (4.0, 1.0), (0.75, 0.05), (2.5, 0.125), (0.5, 1 / 32)]:
        for n in [1, 2, 3, 4, 5, 6, 8, 10, 20, 50]:
            s_, d_ = np.asarray([sigma]), np.asarray([step])
            T = ref._get_half_kernel_size_opt(s_, d_, n)[0]
            Tp = ref._get_half_kernel_size(s_, d_, n)[0]
            a = ref._get_tail_value(s_, d_, n)[0]
            mdw = exp(-3.5 ** 2 / 2)
            csf = _csf_ref(np.asarray([2 * T + 1]), np.asarray([a]), s_, d_, n, mdw)
            csfp = _csf_ref(np.asarray([2 * Tp + 1]), np.asarray([0.0]), s_, d_, n, mdw)
            rows.append([sigma, step, n, T, Tp, a, csf, csfp])
    np.savez_compressed(os.path.join(OUT, 'params.npz'), table=np.asarray(rows, dtype=np.float64))

    # ---- staged small cases --------------------------------------------------------------
    def save_case(name, pts, val, sigma, x0, step, size, n, max_dist=3.5, plain=False):
        st = ref_stages(pts, val, sigma, x0, step, size, n, max_dist, plain)
        dim = len(size)
        full = ref.barnes(pts if dim > 1 else pts.reshape(-1), val, sigma, x0, step, size if dim > 1 else size[0],
                          method='convolution' if plain else 'optimized_convolution', num_iter=n,
                          max_dist=max_dist)
        assert np.array_equal(full.view(np.uint32), st['out32'].view(np.uint32)), name
        np.savez_compressed(os.path.join(OUT, name + '.npz'), pts=pts, val=val,
                            sigma=np.atleast_1d(np.asarray(sigma, dtype=np.float64)),
                            x0=np.atleast_1d(np.asarray(x0, dtype=np.float64)),
                            step=np.atleast_1d(np.asarray(step, dtype=np.float64)),
                            size=np.asarray(size, dtype=np.int64), num_iter=n, max_dist=max_dist,
                            plain=int(plain), **st)
        print(name, 'nan frac %.3f' % np.isnan(st['out32']).mean())

    N = 300
    pts = rng.uniform(-2.0, 22.0, (N, 1))
    val = rng.normal(10, 4, N)
    save_case('case_1d_n4', pts, val, 1.3, 0.0, 0.0625, (512,), 4)
    save_case('case_1d_n6_plain', pts, val, 0.9, 0.0, 0.0625, (512,), 6, plain=True)
    N = 500
    pts = np.stack([rng.uniform(-1.0, 8.0, N), rng.uniform(2.5, 9.5, N)], axis=1)
    # clustered samples -> several samples per cell (ordered accumulation matters)
    pts[:150] = pts[150:300] + rng.normal(0, 0.01, (150, 2))
    val = rng.normal(1000, 12, N)
    save_case('case_2d_n4', pts, val, 0.7, [0.0, 0.0], 0.0625, (192, 144), 4)
    save_case('case_2d_aniso_n3', pts, val, [0.5, 1.0], [0.1, -0.2], [0.05, 0.125], (230, 70), 3)
    save_case('case_2d_n5_plain', pts, val, 0.6, [0.0, 0.0], 0.0625, (192, 144), 5, plain=True)
    save_case('case_2d_T0', pts, val, 0.5, [0.0, 0.0], 1.0, (13, 10), 4)
    N = 400
    pts = rng.uniform(-0.2, 3.0, (N, 3)) * np.asarray([1.0, 0.8, 0.6])
    val = rng.normal(-3, 2, N)
    save_case('case_3d_n4', pts, val, [0.3, 0.28, 0.25], [0.0, 0.0, 0.0], 0.1, (44, 36, 30), 4)
    save_case('case_3d_n2', pts, val, 0.25, [0.0, 0.0, 0.0], 0.1, (44, 36, 30), 2, max_dist=2.5)

    # ---- paper case C1 -------------------------------------------------------------------
    sys.path.insert(0, os.path.join(REF, 'demo'))
    import reader
    pts, val = reader.read_csv_array(os.path.join(REF, 'demo', 'input', 'PressQFF_202007271200_3490.csv'))
    step = 1.0 / 32
    x0 = np.asarray([-26.0 + step, 34.5])
    size = (2400, 1200)
    st = ref_stages(pts, val, 1.0, x0, step, size, 4, 3.5)
    full = ref.barnes(pts, val, 1.0, x0, step, size, num_iter=4)
    assert np.array_equal(full.view(np.uint32), st['out32'].view(np.uint32))
    np.savez_compressed(os.path.join(OUT, 'c1_paper.npz'), pts=pts, val=val, x0=x0, step=step,
                        size=np.asarray(size), sigma=1.0, num_iter=4,
                        sha_out32=sha(st['out32']), sha_out64=sha(st['out64']),
                        sha_vin=sha(st['vin']), sha_win=sha(st['win']),
                        sha_vg=sha(st['vg']), sha_wg=sha(st['wg']),
                        sub_out32=st['out32'][::13, ::17], sub_out64=st['out64'][::13, ::17],
                        offset=st['offset'], T=st['T'], alpha=st['alpha'], csf=st['csf'],
                        nan_frac=np.isnan(st['out32']).mean(),
                        nanmin=np.nanmin(st['out32']), nanmax=np.nanmax(st['out32']))
    print('c1', st['offset'], st['T'], st['alpha'], st['csf'], np.isnan(st['out32']).mean())

    # ---- S2 ------------------------------------------------------------------------------
    proj = refS2.get_lambert_proj()
    lam_pts = reflc.to_map(pts, pts.copy(), *proj)
    for res, full_arrays in ((8, True), (32, False)):
        step = 1.0 / res
        x0 = np.asarray([-26.0 + step, 34.5])
        size = (int(75.0 / step), int(37.5 / step))
        out = refS2.barnes_S2(pts, val, 1.0, x0, step, size, method='optimized_convolution_S2', num_iter=4)
        lam = refS2.barnes_S2(pts, val, 1.0, x0, step, size, method='optimized_convolution_S2', num_iter=4,
                              resample=False)
        d = dict(pts=pts, val=val, x0=x0, step=step, size=np.asarray(size), sigma=1.0, num_iter=4,
                 proj=np.asarray(proj), lam_pts=lam_pts, sha_out=sha(out), sha_lam=sha(lam),
                 sub_out=out[::7, ::11], sub_lam=lam[::7, ::11])
        if full_arrays:
            d['out'] = out
            d['lam'] = lam
        np.savez_compressed(os.path.join(OUT, 's2_res%d.npz' % res), **d)
        print('s2 res', res, out.shape, lam.shape, np.isnan(out).mean())


def exact_cases():
    """ 'naive' / 'radius' (interpolation.py:862-938, :809-855) and 'naive_S2'
    (interpolationS2.py:260-301) through the reference's public API. """
    rng = np.random.default_rng(4242)
    d = {}

    def put(name, fn, pts, val, sigma, x0, step, size, **kw):
        out = fn(pts, val, sigma, np.asarray(x0, dtype=np.float64) if not np.isscalar(x0) else x0, step, size, **kw)
        assert out.dtype == np.float64
        d[name + '_pts'], d[name + '_val'], d[name + '_out'] = pts, val, out
        d[name + '_args'] = np.asarray(np.concatenate([np.atleast_1d(sigma).astype(float).ravel(),
                                                       np.atleast_1d(x0).astype(float).ravel(),
                                                       np.atleast_1d(step).astype(float).ravel(),
                                                       np.atleast_1d(size).astype(float).ravel()]))
        print(name, out.shape, np.isnan(out).mean())

    p2 = rng.uniform(0, 10, (200, 2)); v2 = rng.normal(1000, 10, 200)
    put('naive_2d', ref.barnes, p2, v2, [1.0, 1.0], [0.0, 0.0], [0.25, 0.25], (41, 37), method='naive')
    put('naive_2d_aniso', ref.barnes, p2, v2, [1.0, 0.5], [0.0, 0.5], [0.25, 0.5], (41, 19), method='naive')
    put('radius_2d', ref.barnes, p2, v2, [1.0, 1.0], [0.0, 0.0], [0.25, 0.25], (41, 37), method='radius')
    ps = rng.uniform(0, 3, (20, 2)); vs = rng.normal(0, 1, 20)
    put('radius_2d_sparse', ref.barnes, ps, vs, [0.5, 0.5], [0.0, 0.0], [0.25, 0.25], (41, 37), method='radius')
    put('radius_2d_minw', ref.barnes, ps, vs, [0.5, 0.5], [0.0, 0.0], [0.25, 0.25], (41, 37), method='radius',
        max_dist=2.0, min_weight=0.01)
    p1 = rng.uniform(0, 10, (50, 1)); v1 = rng.normal(0, 1, 50)
    put('naive_1d', ref.barnes, p1, v1, [0.7], [0.0], [0.1], (101,), method='naive')
    p3 = rng.uniform(0, 5, (80, 3)); v3 = rng.normal(0, 1, 80)
    put('naive_3d', ref.barnes, p3, v3, [1.0, 0.8, 0.6], [0.0, 0.0, 0.0], [0.5, 0.5, 0.5], (11, 9, 7), method='naive')
    pl = np.column_stack([rng.uniform(-20, 30, 150), rng.uniform(35, 65, 150)]); vl = rng.normal(1000, 10, 150)
    put('naive_S2', refS2.barnes_S2, pl, vl, [1.5, 1.5], [-20.0, 35.0], [1.0, 1.0], (51, 31), method='naive_S2')

    sys.path.insert(0, os.path.join(REF, 'demo'))
    import reader
    pts, val = reader.read_csv_array(os.path.join(REF, 'demo', 'input', 'PressQFF_202007271200_3490.csv'))
    put('paper_naive', ref.barnes, pts, val, [1.0, 1.0], [-26.0 + 0.5, 34.5], [0.5, 0.5], (150, 75), method='naive')
    put('paper_radius', ref.barnes, pts, val, [1.0, 1.0], [-26.0 + 0.5, 34.5], [0.5, 0.5], (150, 75), method='radius')
    put('paper_naive_S2', refS2.barnes_S2, pts, val, [1.0, 1.0], [-26.0 + 0.5, 34.5], [0.5, 0.5], (150, 75),
        method='naive_S2')
    del d['paper_radius_pts'], d['paper_radius_val'], d['paper_naive_S2_pts'], d['paper_naive_S2_val']
    np.savez_compressed(os.path.join(OUT, 'exact_methods.npz'), **d)


def s2_map_case():
    """ Generalised S2 ("next" row N4): interpolationS2.py:180-202 replayed with another projection
    and map window, using only the reference's functions. """
    rng = np.random.default_rng(777)
    n = 1500
    pts = np.column_stack([rng.uniform(-128.0, -66.0, n), rng.uniform(23.0, 52.0, n)])
    val = 1013.0 + 8.0 * np.sin(pts[:, 0] / 9.0) + 5.0 * np.cos(pts[:, 1] / 5.0) + rng.normal(0, 0.5, n)
    step = np.full(2, 0.25)
    x0 = np.asarray([-125.0, 25.0])
    size = (220, 100)
    sigma = np.full(2, 1.0)
    num_iter = 4
    mdw = exp(-3.5 ** 2 / 2)
    proj = reflc.create_proj(-97.625, 37.375, 29.125, 45.625)
    lons = x0[0] + np.arange(size[0]) * step[0]
    lats = x0[1] + np.arange(size[1]) * step[1]
    border = np.concatenate([np.column_stack([lons, np.full(size[0], lats[0])]),
                             np.column_stack([lons, np.full(size[0], lats[-1])]),
                             np.column_stack([np.full(size[1], lons[0]), lats]),
                             np.column_stack([np.full(size[1], lons[-1]), lats])])
    mapped = reflc.to_map(border, border.copy(), *proj)
    lam_x0 = np.floor(mapped.min(axis=0) - 4.0)
    lam_extent = np.ceil(mapped.max(axis=0) + 4.0) - lam_x0
    lam_size = (int(lam_extent[0] / step[0]), int(lam_extent[1] / step[1]))
    lam_pts = reflc.to_map(pts, pts.copy(), *proj)
    v = val.copy()
    lam = ref._interpolate_opt_convol(lam_pts, v, sigma, lam_x0, step, lam_size, num_iter, mdw)
    out = refS2._resample(lam, lam_x0, x0, step, size, *proj)
    np.savez_compressed(os.path.join(OUT, 's2_map_na.npz'), pts=pts, val=val, x0=x0, step=step,
                        size=np.asarray(size), sigma=sigma, num_iter=num_iter, proj=np.asarray(proj),
                        lam_x0=lam_x0, lam_extent=lam_extent, lam_pts=lam_pts, lam=lam, out=out)
    print('s2 map', lam.shape, out.shape, lam_x0, lam_extent, np.isnan(out).mean(), np.isnan(lam).mean())


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'exact':
        exact_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == 's2map':
        s2_map_case()
    else:
        main()
        exact_cases()
        s2_map_case()
