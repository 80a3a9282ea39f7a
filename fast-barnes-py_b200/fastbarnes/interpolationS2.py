# -*- coding: utf-8 -*-
"""
B200-native drop-in for `fastbarnes.interpolationS2` of MeteoSwiss/fast-barnes-py (v2.0.0),
method 'optimized_convolution_S2': the samples are projected to a fixed Lambert conformal map,
the Euclidean optimized convolution runs on the fixed Lambert grid (lam_x0 = (-32, -2),
64 x 44 degrees), and the Lambert field is bilinearly resampled to the requested lon/lat grid
(reference interpolationS2.py:144-254); method 'naive_S2' is the exact Gaussian sum over
spherical distances (:260-301).  All steps are CUDA kernels behind include/fastbarnes_b200.h;
there is no CPU fallback.
"""
from math import exp

import numpy as np

from . import _lib
from .interpolation import _per_axis, _grid_size, _run_exact
from .util import lambert_conformal

__all__ = ['barnes_S2', 'interpolate_opt_convol_S2_part1', 'interpolate_opt_convol_S2_part2', 'get_lambert_proj']


def barnes_S2(pts, val, sigma, x0, step, size, method='optimized_convolution', num_iter=4, max_dist=3.5,
              resample=True):
    """
    Barnes interpolation on the sphere S^2 (spherical distances in degrees) for sample points
    `pts` (N, 2) given as lon/lat.  Signature and result as the reference's `barnes_S2`
    (interpolationS2.py:32-138): float32 array of shape (size[1], size[0]), or the Lambert-grid
    field of shape (int(44/step), int(64/step)) if `resample` is False.

    Accepted methods: 'optimized_convolution_S2' and 'naive_S2' (the exact Gaussian sum over
    spherical distances, float64 result; agrees with the reference to rounding).  The reference's
    default string 'optimized_convolution' is not accepted by the reference itself (it raises
    RuntimeError); here it is taken as an alias of 'optimized_convolution_S2'.
    """
    dim = pts.shape[1]
    sigma = _per_axis('sigma', sigma, dim)
    x0 = _per_axis('x0', x0, dim)
    step = _per_axis('step', step, dim)
    size = _grid_size(size, dim)
    max_dist_weight = exp(-max_dist ** 2 / 2)

    if method in ('optimized_convolution_S2', 'optimized_convolution'):
        return _interpolate_opt_convol_S2(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, resample)
    if method == 'naive_S2':
        pts_c, val_c = _samples(pts, val)
        return _run_exact(pts_c, val_c, sigma, x0, step, size, _lib.METHOD_NAIVE_S2, max_dist_weight, 0.0)
    raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))


def _samples(pts, val):
    pts_c = np.ascontiguousarray(pts, dtype=np.float64)
    val_c = np.ascontiguousarray(val, dtype=np.float64)
    if pts_c.ndim != 2 or pts_c.shape[1] != 2 or val_c.shape != (pts_c.shape[0],):
        raise RuntimeError('expected pts of shape (N, 2) and val of shape (N)')
    if pts_c.shape[0] == 0:
        raise ValueError('zero-size array to reduction operation minimum which has no identity')
    return pts_c, val_c


def _interpolate_opt_convol_S2(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, resample):
    """ Reference interpolationS2.py:144-177; with resample the Lambert field never leaves the GPU. """
    if not resample:
        return interpolate_opt_convol_S2_part1(pts, val, sigma, x0, step, size, num_iter, max_dist_weight)[0]
    pts_c, val_c = _samples(pts, val)
    proj = np.asarray(get_lambert_proj(), dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    sz = np.asarray(size, dtype=np.int64)
    res = np.empty((int(size[1]), int(size[0])), dtype=np.float32)
    rc = _lib.lib().fb_barnes_s2_host(pts_c.shape[0], _lib.dptr(pts_c), _lib.dptr(val_c), _lib.dptr(sigma),
                                      _lib.dptr(x0), _lib.dptr(step), sz.ctypes.data_as(_lib.c_i64_p),
                                      int(num_iter), float(max_dist_weight), _lib.dptr(proj),
                                      res.ctypes.data_as(_lib.c_float_p))
    _lib.check(rc)
    return res


def interpolate_opt_convol_S2_part1(pts, val, sigma, x0, step, size, num_iter, max_dist_weight):
    """
    The convolution part in Lambert space (reference interpolationS2.py:180-196).
    Returns (lam_field, lam_x0, x0, step, size, lambert_proj) like the reference.
    """
    pts_c, val_c = _samples(pts, val)
    lambert_proj = get_lambert_proj()
    proj = np.asarray(lambert_proj, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    lam_x0 = np.asarray([-32.0, -2.0])
    lam_size = (int(64.0 / step[0]), int(44.0 / step[1]))
    lam_field = np.empty((lam_size[1], lam_size[0]), dtype=np.float32)
    rc = _lib.lib().fb_s2_part1_host(pts_c.shape[0], _lib.dptr(pts_c), _lib.dptr(val_c), _lib.dptr(sigma),
                                     _lib.dptr(step), int(num_iter), float(max_dist_weight), _lib.dptr(proj),
                                     lam_field.ctypes.data_as(_lib.c_float_p))
    _lib.check(rc)
    return (lam_field, lam_x0, x0, step, size, lambert_proj)


def interpolate_opt_convol_S2_part2(lam_field, lam_x0, x0, step, size, lambert_proj):
    """ The back-projection part (reference interpolationS2.py:199-202). """
    return _resample(lam_field, lam_x0, x0, step, size, *lambert_proj)


def get_lambert_proj():
    """ The Lambert projection of the test example (reference interpolationS2.py:205-208). """
    return lambert_conformal.create_proj(11.5, 34.5, 42.5, 65.5)


def _resample(lam_field, lam_x0, x0, step, size, center_lon, n, n_inv, F, rho0):
    """ Resamples the Lambert grid field to the lon/lat grid (reference interpolationS2.py:211-254). """
    lam = np.ascontiguousarray(lam_field, dtype=np.float32)
    proj = np.asarray([center_lon, n, n_inv, F, rho0], dtype=np.float64)
    lam_x0 = np.ascontiguousarray(lam_x0, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    sz = np.asarray(size, dtype=np.int64)
    res = np.empty((int(size[1]), int(size[0])), dtype=np.float32)
    rc = _lib.lib().fb_s2_resample_host(lam.ctypes.data_as(_lib.c_float_p), lam.shape[1], lam.shape[0],
                                        _lib.dptr(lam_x0), _lib.dptr(x0), _lib.dptr(step),
                                        sz.ctypes.data_as(_lib.c_i64_p), _lib.dptr(proj),
                                        res.ctypes.data_as(_lib.c_float_p))
    _lib.check(rc)
    return res
