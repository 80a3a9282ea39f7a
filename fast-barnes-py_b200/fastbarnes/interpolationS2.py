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

__all__ = ['barnes_S2', 'interpolate_opt_convol_S2_part1', 'interpolate_opt_convol_S2_part2', 'get_lambert_proj',
           'LambertMap']


class LambertMap(object):
    """
    The Lambert conformal map the S2 convolution runs on: projection constants `proj` =
    (center_lon, n, n_inv, F, rho0) as returned by `lambert_conformal.create_proj`, and the window
    of the map grid in map coordinates (degrees): start `lam_x0` and `lam_extent`; the grid has
    `int(lam_extent / step)` points per axis.  The reference hard-codes one such map for Europe
    (interpolationS2.py:187-188, :208) -- `LambertMap.default()`; `LambertMap.for_grid()` derives one
    for any lon/lat target grid ("next" row N4 of SURVEY section 8f).
    """

    def __init__(self, proj, lam_x0, lam_extent):
        self.proj = tuple(float(p) for p in proj)
        self.lam_x0 = (float(lam_x0[0]), float(lam_x0[1]))
        self.lam_extent = (float(lam_extent[0]), float(lam_extent[1]))
        if len(self.proj) != 5 or not (self.lam_extent[0] > 0.0 and self.lam_extent[1] > 0.0):
            raise RuntimeError('invalid Lambert map: ' + repr((proj, lam_x0, lam_extent)))

    @classmethod
    def default(cls):
        """ The reference's fixed map: centre (11.5, 34.5), standard parallels 42.5 / 65.5,
        window start (-32, -2), extent 64 x 44 map degrees. """
        return cls(get_lambert_proj(), (-32.0, -2.0), (64.0, 44.0))

    @classmethod
    def for_grid(cls, x0, step, size, margin=0.0):
        """
        A map for the lon/lat grid (x0, step, size): projection centred on the grid, standard
        parallels at 1/6 and 5/6 of its latitude range, window = bounding box of the projected grid
        border widened by `margin` map degrees on every side (samples beyond the target grid that
        should still contribute, e.g. max_dist*sigma).
        """
        x0 = _per_axis('x0', x0, 2)
        step = _per_axis('step', step, 2)
        size = _grid_size(size, 2)
        lon0, lon1 = x0[0], x0[0] + (size[0] - 1) * step[0]
        lat0, lat1 = x0[1], x0[1] + (size[1] - 1) * step[1]
        if not (-89.0 < min(lat0, lat1) and max(lat0, lat1) < 89.0 and (lat0 > 0.0) == (lat1 > 0.0)):
            raise RuntimeError('a Lambert conformal map needs a grid within one hemisphere, away from the poles')
        proj = lambert_conformal.create_proj(0.5 * (lon0 + lon1), 0.5 * (lat0 + lat1),
                                             lat0 + (lat1 - lat0) / 6.0, lat1 - (lat1 - lat0) / 6.0)
        lons = x0[0] + np.arange(size[0]) * step[0]
        lats = x0[1] + np.arange(size[1]) * step[1]
        border = np.concatenate([np.column_stack([lons, np.full(size[0], lat0)]),
                                 np.column_stack([lons, np.full(size[0], lat1)]),
                                 np.column_stack([np.full(size[1], lon0), lats]),
                                 np.column_stack([np.full(size[1], lon1), lats])])
        mapped = lambert_conformal.to_map(border, np.empty_like(border), *proj)
        lo = mapped.min(axis=0) - margin - step
        hi = mapped.max(axis=0) + margin + 2 * step
        return cls(proj, lo, hi - lo)

    def lam_size(self, step):
        return (int(self.lam_extent[0] / step[0]), int(self.lam_extent[1] / step[1]))

    def _struct(self):
        m = _lib.FbS2Map()
        for i in range(5):
            m.proj[i] = self.proj[i]
        for i in range(2):
            m.lam_x0[i] = self.lam_x0[i]
            m.lam_extent[i] = self.lam_extent[i]
        return m


def _resolve_map(lambert_map, x0, step, size, margin):
    if lambert_map is None:
        return LambertMap.default()
    if isinstance(lambert_map, str):
        if lambert_map != 'auto':
            raise RuntimeError("lambert_map should be None, 'auto' or a LambertMap: " + lambert_map)
        return LambertMap.for_grid(x0, step, size, margin)
    if not isinstance(lambert_map, LambertMap):
        raise RuntimeError("lambert_map should be None, 'auto' or a LambertMap")
    return lambert_map


def barnes_S2(pts, val, sigma, x0, step, size, method='optimized_convolution', num_iter=4, max_dist=3.5,
              resample=True, *, lambert_map=None):
    """
    Barnes interpolation on the sphere S^2 (spherical distances in degrees) for sample points
    `pts` (N, 2) given as lon/lat.  Signature and result as the reference's `barnes_S2`
    (interpolationS2.py:32-138): float32 array of shape (size[1], size[0]), or the Lambert-grid
    field of shape (int(44/step), int(64/step)) if `resample` is False.

    Accepted methods: 'optimized_convolution_S2' and 'naive_S2' (the exact Gaussian sum over
    spherical distances, float64 result; agrees with the reference to rounding).  The reference's
    default string 'optimized_convolution' is not accepted by the reference itself (it raises
    RuntimeError); here it is taken as an alias of 'optimized_convolution_S2'.

    lambert_map (extension, 'optimized_convolution_S2' only): None = the reference's fixed European
    map; 'auto' = `LambertMap.for_grid(x0, step, size, margin=max_dist*max(sigma))`; or a LambertMap.
    Output pixels whose bilinear stencil leaves the map window are NaN.
    """
    dim = pts.shape[1]
    sigma = _per_axis('sigma', sigma, dim)
    x0 = _per_axis('x0', x0, dim)
    step = _per_axis('step', step, dim)
    size = _grid_size(size, dim)
    max_dist_weight = exp(-max_dist ** 2 / 2)

    if method in ('optimized_convolution_S2', 'optimized_convolution'):
        lmap = _resolve_map(lambert_map, x0, step, size, max_dist * float(np.max(sigma)))
        return _interpolate_opt_convol_S2(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, resample, lmap)
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


def _interpolate_opt_convol_S2(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, resample, lambert_map=None):
    """ Reference interpolationS2.py:144-177; with resample the Lambert field never leaves the GPU. """
    lmap = lambert_map if lambert_map is not None else LambertMap.default()
    if not resample:
        return interpolate_opt_convol_S2_part1(pts, val, sigma, x0, step, size, num_iter, max_dist_weight,
                                               lambert_map=lmap)[0]
    pts_c, val_c = _samples(pts, val)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    sz = np.asarray(size, dtype=np.int64)
    res = np.empty((int(size[1]), int(size[0])), dtype=np.float32)
    rc = _lib.lib().fb_barnes_s2_map_host(pts_c.shape[0], _lib.dptr(pts_c), _lib.dptr(val_c), _lib.dptr(sigma),
                                          _lib.dptr(x0), _lib.dptr(step), sz.ctypes.data_as(_lib.c_i64_p),
                                          int(num_iter), float(max_dist_weight), lmap._struct(),
                                          res.ctypes.data_as(_lib.c_float_p))
    _lib.check(rc)
    return res


def interpolate_opt_convol_S2_part1(pts, val, sigma, x0, step, size, num_iter, max_dist_weight, *, lambert_map=None):
    """
    The convolution part in Lambert space (reference interpolationS2.py:180-196).
    Returns (lam_field, lam_x0, x0, step, size, lambert_proj) like the reference.
    """
    pts_c, val_c = _samples(pts, val)
    lmap = lambert_map if lambert_map is not None else LambertMap.default()
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    lam_x0 = np.asarray(lmap.lam_x0)
    lam_size = lmap.lam_size(step)
    lam_field = np.empty((lam_size[1], lam_size[0]), dtype=np.float32)
    rc = _lib.lib().fb_s2_part1_map_host(pts_c.shape[0], _lib.dptr(pts_c), _lib.dptr(val_c), _lib.dptr(sigma),
                                         _lib.dptr(step), int(num_iter), float(max_dist_weight), lmap._struct(),
                                         lam_field.ctypes.data_as(_lib.c_float_p))
    _lib.check(rc)
    return (lam_field, lam_x0, x0, step, size, lmap.proj)


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
