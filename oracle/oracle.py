# -*- coding: utf-8 -*-
"""
oracle.py -- ctypes front end of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Wraps oracle/fb_oracle.c, the plain-C restatement of the reference's
optimized-convolution Barnes interpolation.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` leg may import this module.

Function names mirror the reference (fastbarnes/interpolation.py,
fastbarnes/interpolationS2.py, fastbarnes/util/lambert_conformal.py); each
docstring cites the reference file:line.  Parity pin: see the header of
fb_oracle.c and tests/test_oracle_vs_reference.py / tests/golden/.
"""
import ctypes
import os
import subprocess
from math import exp

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libfb_oracle.so')

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)
_c_i32_p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """ Compiles oracle/fb_oracle.c into oracle/_build/ (make). """
    src = os.path.join(_HERE, 'fb_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.fbo_normalize_values.restype = ctypes.c_double
        L.fbo_normalize_values.argtypes = [_c_double_p, ctypes.c_int64]
        L.fbo_inject_data.restype = None
        L.fbo_inject_data.argtypes = [ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                      ctypes.c_int64, _c_double_p, _c_double_p, _c_i64_p]
        L.fbo_half_kernel_size_opt.restype = ctypes.c_int32
        L.fbo_half_kernel_size_opt.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.fbo_half_kernel_size.restype = ctypes.c_int32
        L.fbo_half_kernel_size.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.fbo_tail_value.restype = ctypes.c_double
        L.fbo_tail_value.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.fbo_conv_scale_factor.restype = ctypes.c_double
        L.fbo_conv_scale_factor.argtypes = [ctypes.c_int, _c_i32_p, _c_double_p, _c_double_p, _c_double_p,
                                            ctypes.c_int, ctypes.c_double]
        L.fbo_accumulate_tail_array.restype = ctypes.c_int
        L.fbo_accumulate_tail_array.argtypes = [_c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_double]
        L.fbo_accumulate_array.restype = ctypes.c_int
        L.fbo_accumulate_array.argtypes = [_c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int]
        L.fbo_convolve.restype = None
        L.fbo_convolve.argtypes = [ctypes.c_int, _c_double_p, _c_double_p, _c_i64_p, _c_i32_p, ctypes.c_int,
                                   _c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.fbo_interpolate.restype = ctypes.c_double
        L.fbo_interpolate.argtypes = [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int64, _c_double_p,
                                      _c_double_p, _c_double_p, _c_i64_p, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_int, _c_float_p, _c_double_p, _c_double_p,
                                      _c_double_p, _c_double_p, _c_double_p]
        L.fbo_lambert_create_proj.restype = None
        L.fbo_lambert_create_proj.argtypes = [ctypes.c_double] * 4 + [_c_double_p]
        L.fbo_lambert_to_map.restype = None
        L.fbo_lambert_to_map.argtypes = [_c_double_p, _c_double_p, ctypes.c_int64, _c_double_p]
        L.fbo_resample.restype = None
        L.fbo_resample.argtypes = [_c_float_p, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p,
                                   _c_i64_p, _c_double_p, _c_float_p, ctypes.c_int]
        L.fbo_interpolate_exact.restype = ctypes.c_double
        L.fbo_interpolate_exact.argtypes = [ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int64,
                                            _c_double_p, _c_double_p, _c_double_p, _c_i64_p, ctypes.c_double,
                                            ctypes.c_double, _c_double_p, ctypes.c_int]
        L.fbo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _f64(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a


def _vec(v, dim):
    if isinstance(v, (list, tuple, np.ndarray)):
        return np.ascontiguousarray(np.asarray(v, dtype=np.float64))
    return np.full(dim, v, dtype=np.float64)


def max_threads():
    return lib().fbo_max_threads()


# ---------------------------------------------------------------------------

def _normalize_values(val):
    """ fastbarnes/interpolation.py:205-212. In place on float64 `val`; returns offset. """
    assert val.dtype == np.float64 and val.flags.c_contiguous
    return lib().fbo_normalize_values(_dp(val), val.shape[0])


def _inject_data(vg, wg, pts, val, x0, step, size):
    """ fastbarnes/interpolation.py:219-322 (_inject_data_1d/_2d/_3d by len(size)). """
    dim = len(size)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, dim)
    sz = np.asarray(size, dtype=np.int64)
    lib().fbo_inject_data(dim, _dp(vg), _dp(wg), _dp(pts), _dp(val), pts.shape[0], _dp(_vec(x0, dim)),
                          _dp(_vec(step, dim)), sz.ctypes.data_as(_c_i64_p))


def _get_half_kernel_size_opt(sigma, step, num_iter):
    """ fastbarnes/interpolation.py:549-552 (array version). """
    sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
    step = np.atleast_1d(np.asarray(step, dtype=np.float64))
    return np.asarray([lib().fbo_half_kernel_size_opt(s, d, int(num_iter)) for s, d in zip(sigma, step)],
                      dtype=np.int32)


def _get_half_kernel_size(sigma, step, num_iter):
    """ fastbarnes/interpolation.py:783-785 (array version). """
    sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
    step = np.atleast_1d(np.asarray(step, dtype=np.float64))
    return np.asarray([lib().fbo_half_kernel_size(s, d, int(num_iter)) for s, d in zip(sigma, step)],
                      dtype=np.int32)


def _get_tail_value(sigma, step, num_iter):
    """ fastbarnes/interpolation.py:561-569 (array version). """
    sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
    step = np.atleast_1d(np.asarray(step, dtype=np.float64))
    return np.asarray([lib().fbo_tail_value(s, d, int(num_iter)) for s, d in zip(sigma, step)],
                      dtype=np.float64)


def conv_scale_factor(kernel_size, tail_value, sigma, step, num_iter, max_dist_weight):
    """ fastbarnes/interpolation.py:424-425. """
    ks = np.ascontiguousarray(kernel_size, dtype=np.int32)
    dim = len(ks)
    tv = _vec(tail_value, dim)
    return lib().fbo_conv_scale_factor(dim, ks.ctypes.data_as(_c_i32_p), _dp(tv), _dp(_vec(sigma, dim)),
                                       _dp(_vec(step, dim)), int(num_iter), float(max_dist_weight))


def _accumulate_tail_array(in_arr, h_arr, arr_len, rect_len, num_iter, alpha):
    """ fastbarnes/interpolation.py:485-533. Returns the array that holds the result. """
    which = lib().fbo_accumulate_tail_array(_dp(in_arr), _dp(h_arr), int(arr_len), int(rect_len),
                                            int(num_iter), float(alpha))
    return h_arr if which else in_arr


def _accumulate_array(in_arr, h_arr, arr_len, rect_len, num_iter):
    """ fastbarnes/interpolation.py:729-772. Returns the array that holds the result. """
    which = lib().fbo_accumulate_array(_dp(in_arr), _dp(h_arr), int(arr_len), int(rect_len), int(num_iter))
    return h_arr if which else in_arr


def _convolve_tail(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight,
                   plain=False, nthreads=1):
    """
    fastbarnes/interpolation.py:373-479 (_convolve_tail_1d/_2d/_3d by len(size));
    plain=True: :617-724 (_convolve_1d/_2d/_3d). In place on vg, wg.
    """
    dim = len(size)
    ks = np.ascontiguousarray(kernel_size, dtype=np.int32)
    tv = np.zeros(dim) if plain else _vec(tail_value, dim)
    csf = conv_scale_factor(ks, tv, sigma, step, num_iter, max_dist_weight)
    sz = np.asarray(size, dtype=np.int64)
    lib().fbo_convolve(dim, _dp(vg), _dp(wg), sz.ctypes.data_as(_c_i64_p), ks.ctypes.data_as(_c_i32_p),
                       int(num_iter), _dp(tv), int(plain), csf, int(nthreads))


def _interpolate_opt_convol(pts, val, sigma, x0, step, size, num_iter, max_dist_weight,
                            plain=False, nthreads=1, stages=False):
    """
    fastbarnes/interpolation.py:329-367 (plain=True: :575-612). `val` is modified in
    place like in the reference.  With stages=True returns a dict holding the
    intermediate fields as well (vin/win post-injection, vg/wg post-sweep, out64
    pre-cast quotient, out32, offset).
    """
    dim = len(size)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, dim)
    assert val.dtype == np.float64 and val.flags.c_contiguous
    sz = np.asarray(size, dtype=np.int64)
    rsize = tuple(int(s) for s in size[::-1])
    out32 = np.empty(rsize, dtype=np.float32)
    opt = {}
    if stages:
        for name in ('out64', 'vg', 'wg', 'vin', 'win'):
            opt[name] = np.empty(rsize, dtype=np.float64)
    null = ctypes.cast(None, _c_double_p)
    offset = lib().fbo_interpolate(
        dim, _dp(pts), _dp(val), pts.shape[0], _dp(_vec(sigma, dim)), _dp(_vec(x0, dim)),
        _dp(_vec(step, dim)), sz.ctypes.data_as(_c_i64_p), int(num_iter), float(max_dist_weight),
        int(plain), int(nthreads), out32.ctypes.data_as(_c_float_p),
        _dp(opt['out64']) if stages else null, _dp(opt['vg']) if stages else null,
        _dp(opt['wg']) if stages else null, _dp(opt['vin']) if stages else null,
        _dp(opt['win']) if stages else null)
    if stages:
        opt['out32'] = out32
        opt['offset'] = offset
        return opt
    return out32


def _interpolate_exact(kind, pts, val, sigma, x0, step, size, max_dist_weight=0.0, min_weight=0.001, nthreads=1):
    """ naive (kind 0, interpolation.py:862-938), naive_S2 (kind 1, interpolationS2.py:260-301) and
    radius (kind 2, interpolation.py:809-855) summed in sample order; val is centred in place. """
    dim = len(size)
    pts = _f64(pts)
    out = np.empty(tuple(size)[::-1], dtype=np.float64)
    sz = np.asarray(size, dtype=np.int64)
    lib().fbo_interpolate_exact(int(kind), dim, _dp(pts), _dp(val), len(val), _dp(_f64(sigma)), _dp(_f64(x0)),
                                _dp(_f64(step)), sz.ctypes.data_as(_c_i64_p), float(max_dist_weight),
                                float(min_weight), _dp(out), int(nthreads))
    return out


def barnes(pts, val, sigma, x0, step, size, method='optimized_convolution', num_iter=4, max_dist=3.5,
           min_weight=0.001, nthreads=1):
    """ fastbarnes/interpolation.py:31-199 (all four methods; 'radius' by exhaustive search). """
    pts = np.asarray(pts, dtype=np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    dim = pts.shape[1]
    val = np.array(val, dtype=np.float64, copy=True)
    if not isinstance(size, (list, tuple, np.ndarray)):
        size = (size,)
    size = tuple(int(s) for s in size)
    max_dist_weight = exp(-max_dist ** 2 / 2)
    if method == 'naive':
        return _interpolate_exact(0, pts, val, _vec(sigma, dim), _vec(x0, dim), _vec(step, dim), size,
                                  nthreads=nthreads)
    if method == 'radius':
        sg = _vec(sigma, dim)
        if dim != 2:
            raise RuntimeError('radius algorithm works only in 2D but data is: ' + str(dim) + 'D')
        if sg[0] != sg[1]:
            raise RuntimeError('radius algorithm in 2D works only for scalar sigma value but sigma is: ' + str(sg))
        return _interpolate_exact(2, pts, val, sg, _vec(x0, dim), _vec(step, dim), size, max_dist_weight,
                                  min_weight, nthreads=nthreads)
    if method not in ('optimized_convolution', 'convolution'):
        raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))
    return _interpolate_opt_convol(pts, val, _vec(sigma, dim), _vec(x0, dim), _vec(step, dim), size,
                                   num_iter, max_dist_weight, plain=(method == 'convolution'),
                                   nthreads=nthreads)


# ---------------------------------------------------------------------------
# S2 path

def create_proj(center_lon, center_lat, lat1, lat2):
    """ fastbarnes/util/lambert_conformal.py:50-94. Returns (center_lon, n, n_inv, F, rho0). """
    proj = np.empty(5, dtype=np.float64)
    lib().fbo_lambert_create_proj(center_lon, center_lat, lat1, lat2, _dp(proj))
    return tuple(float(p) for p in proj)


def get_lambert_proj():
    """ fastbarnes/interpolationS2.py:205-208. """
    return create_proj(11.5, 34.5, 42.5, 65.5)


def to_map(geoc, mapc, center_lon, n, n_inv, F, rho0):
    """ fastbarnes/util/lambert_conformal.py:113-123. """
    proj = np.asarray([center_lon, n, n_inv, F, rho0], dtype=np.float64)
    geoc = np.ascontiguousarray(geoc, dtype=np.float64)
    lib().fbo_lambert_to_map(_dp(geoc), _dp(mapc), geoc.shape[0], _dp(proj))
    return mapc


def _resample(lam_field, lam_x0, x0, step, size, center_lon, n, n_inv, F, rho0, nthreads=1):
    """ fastbarnes/interpolationS2.py:212-254. """
    proj = np.asarray([center_lon, n, n_inv, F, rho0], dtype=np.float64)
    lam_field = np.ascontiguousarray(lam_field, dtype=np.float32)
    sz = np.asarray(size, dtype=np.int64)
    res = np.empty((int(size[1]), int(size[0])), dtype=np.float32)
    lib().fbo_resample(lam_field.ctypes.data_as(_c_float_p), lam_field.shape[1], _dp(_vec(lam_x0, 2)),
                       _dp(_vec(x0, 2)), _dp(_vec(step, 2)), sz.ctypes.data_as(_c_i64_p), _dp(proj),
                       res.ctypes.data_as(_c_float_p), int(nthreads))
    return res


def interpolate_opt_convol_S2_part1(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, nthreads=1,
                                    lambert_map=None):
    """ fastbarnes/interpolationS2.py:180-196.  val is centred in place like in the reference.
    lambert_map = (proj, lam_x0, lam_extent) replaces the hard-coded map of :187-188, :208. """
    if lambert_map is None:
        lambert_proj, lam_x0, lam_extent = get_lambert_proj(), (-32.0, -2.0), (64.0, 44.0)
    else:
        lambert_proj, lam_x0, lam_extent = lambert_map
        lambert_proj = tuple(float(p) for p in lambert_proj)
    lam_x0 = np.asarray(lam_x0, dtype=np.float64)
    lam_size = (int(lam_extent[0] / step[0]), int(lam_extent[1] / step[1]))
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    lam_pts = to_map(pts, pts.copy(), *lambert_proj)
    lam_field = _interpolate_opt_convol(lam_pts, val, sigma, lam_x0, step, lam_size, num_iter,
                                        max_dist_weight, nthreads=nthreads)
    return (lam_field, lam_x0, x0, step, size, lambert_proj)


def interpolate_opt_convol_S2_part2(lam_field, lam_x0, x0, step, size, lambert_proj, nthreads=1):
    """ fastbarnes/interpolationS2.py:199-202. """
    return _resample(lam_field, lam_x0, x0, step, size, *lambert_proj, nthreads=nthreads)


def barnes_S2(pts, val, sigma, x0, step, size, method='optimized_convolution_S2', num_iter=4, max_dist=3.5,
              resample=True, nthreads=1, lambert_map=None):
    """ fastbarnes/interpolationS2.py:32-138. """
    val = np.array(val, dtype=np.float64, copy=True)
    if method == 'naive_S2':
        return _interpolate_exact(1, np.asarray(pts, dtype=np.float64), val, _vec(sigma, 2), _vec(x0, 2),
                                  _vec(step, 2), tuple(int(s) for s in size), nthreads=nthreads)
    if method != 'optimized_convolution_S2':
        raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))
    res1 = interpolate_opt_convol_S2_part1(pts, val, _vec(sigma, 2), _vec(x0, 2), _vec(step, 2),
                                           tuple(int(s) for s in size), num_iter,
                                           exp(-max_dist ** 2 / 2), nthreads=nthreads, lambert_map=lambert_map)
    if resample:
        return interpolate_opt_convol_S2_part2(*res1, nthreads=nthreads)
    return res1[0]
