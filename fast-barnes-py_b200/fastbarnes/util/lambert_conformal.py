# -*- coding: utf-8 -*-
"""
Lambert conformal conic projection, forward direction only (what the S2 hot path needs).
Mirrors `fastbarnes.util.lambert_conformal` of the reference: a projection 'instance' is the
tuple (center_lon, n, n_inv, F, rho0).  create_proj runs on the host (five scalars, libm like
the reference); to_map runs as a CUDA kernel.  The inverse mapping (to_geo / to_geo2) and
get_scale are only used by the reference's plotting demos and are out of scope.
"""
import numpy as np

from .. import _lib


def create_proj(center_lon, center_lat, lat1, lat2):
    """ Reference util/lambert_conformal.py:50-94. Returns (center_lon, n, n_inv, F, rho0). """
    proj = np.empty(5, dtype=np.float64)
    _lib.check(_lib.lib().fb_lambert_create_proj(float(center_lon), float(center_lat), float(lat1),
                                                 float(lat2), _lib.dptr(proj)))
    return tuple(float(p) for p in proj)


def to_map(geoc, mapc, center_lon, n, n_inv, F, rho0):
    """
    Reference util/lambert_conformal.py:113-123: maps the lon/lat coordinates `geoc` (N, 2) to
    Lambert map coordinates, stores them in the preallocated `mapc` (N, 2) and returns it.
    """
    proj = np.asarray([center_lon, n, n_inv, F, rho0], dtype=np.float64)
    src = np.ascontiguousarray(geoc, dtype=np.float64)
    dst = mapc if (mapc.flags.c_contiguous and mapc.dtype == np.float64) else np.empty(src.shape, np.float64)
    _lib.check(_lib.lib().fb_lambert_to_map_host(_lib.dptr(src), _lib.dptr(dst), src.shape[0], _lib.dptr(proj)))
    if dst is not mapc:
        mapc[...] = dst
    return mapc
