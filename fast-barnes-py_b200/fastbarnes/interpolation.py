# -*- coding: utf-8 -*-
"""
B200-native drop-in for `fastbarnes.interpolation` of MeteoSwiss/fast-barnes-py (v2.0.0),
restricted to the convolution path: `barnes(..., method='optimized_convolution')` (and its
sibling 'convolution', which is the same kernel with tail value 0).

Same names, argument meaning and error behaviour as the reference
(fastbarnes/interpolation.py); the arithmetic runs in hand-written sm_100a CUDA kernels
behind the C ABI of include/fastbarnes_b200.h and is bit-identical to the reference's
Numba loops (same fp64 operation order, no FMA).  There is no CPU fallback.

Extensions (keyword-only / new functions, not present in the reference):
  barnes(..., return_float64=True)   additionally returns the pre-cast fp64 quotient
  barnes_batched(...)                many independent fields (time steps, ensemble members)
                                     on one grid in a single call
  BarnesDevice                       device-resident interface on torch CUDA tensors
"""
from math import exp

import numpy as np

from . import _lib

__all__ = ['barnes', 'barnes_batched', 'BarnesDevice', 'get_half_kernel_size_opt', 'get_tail_value',
           'get_half_kernel_size', 'get_sigma_effective']

_CONV_METHODS = {'optimized_convolution': _lib.METHOD_OPTIMIZED_CONVOLUTION,
                 'convolution': _lib.METHOD_CONVOLUTION}


# ---------------------------------------------------------------------------------------------
# argument normalisation (reference: interpolation.py:104-164)

def _per_axis(name, value, dim):
    """ scalar -> length-dim float64 vector; sequences must have length dim (:129-153). """
    if isinstance(value, (list, tuple, np.ndarray)):
        if len(value) != dim:
            raise RuntimeError('specified ' + name + ' with invalid length: ' + str(len(value)))
        return np.asarray(value, dtype=np.float64)
    return np.full(dim, value, dtype=np.float64)


def _grid_size(size, dim):
    """ (:156-163) """
    if isinstance(size, (list, tuple, np.ndarray)):
        if len(size) != dim:
            raise RuntimeError('specified size with invalid length: ' + str(len(size)))
        return tuple(int(s) for s in size)
    if dim != 1:
        raise RuntimeError('array size should be array-like of length: ' + str(dim))
    return (int(size),)


def _check_samples(pts, val):
    """ (:104-123) returns pts as an (N, M) array. """
    if not isinstance(pts, np.ndarray):
        raise RuntimeError('specified pts is not a numpy ndarray')
    if pts.ndim > 2:
        raise RuntimeError('expected pts array of shape (N, M) but was: ' + str(pts.shape))
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    dim = pts.shape[1]
    if dim < 1 or dim > 3:
        raise RuntimeError('Barnes interpolation supports only sample points in dimensions 1, 2 or 3')
    if not isinstance(val, np.ndarray):
        raise RuntimeError('specified val is not a numpy ndarray')
    if val.ndim > 1:
        raise RuntimeError('expected val array of shape (N) but was: ' + str(val.shape))
    if val.shape[0] != pts.shape[0]:
        raise RuntimeError('pts and val arrays have inconsistent shapes: ' + str(pts.shape) + ' vs. ' + str(val.shape))
    return pts


FLAG_SEGMENTED_1D = 1
FLAG_FP32 = 2


def _precision_flag(precision, dim, return_float64=False):
    if precision == 'fp64':
        return 0
    if precision != 'fp32':
        raise RuntimeError("precision should be 'fp64' or 'fp32': " + str(precision))
    if dim < 2:
        raise RuntimeError('fp32 working precision covers 2D and 3D grids only')
    if return_float64:
        raise RuntimeError('fp32 working precision has no float64 quotient')
    return FLAG_FP32


def _problem(dim, sigma, x0, step, size, method_id, num_iter, max_dist_weight, nfields=1, flags=0):
    p = _lib.FbProblem()
    p.dim = dim
    p.method = method_id
    p.num_iter = int(num_iter)
    p.flags = int(flags)
    p.nfields = int(nfields)
    for m in range(3):
        p.size[m] = int(size[m]) if m < dim else 1
        p.sigma[m] = float(sigma[m]) if m < dim else 1.0
        p.x0[m] = float(x0[m]) if m < dim else 0.0
        p.step[m] = float(step[m]) if m < dim else 1.0
    p.max_dist_weight = float(max_dist_weight)
    return p


def _check_kernel_vs_grid(method, sigma, step, size, num_iter):
    """ (:171-175, :180-184) the rectangular kernel must be smaller than the grid. """
    if method == 'optimized_convolution':
        kernel_size = 2 * _get_half_kernel_size_opt(sigma, step, num_iter) + 1
    else:
        kernel_size = 2 * _get_half_kernel_size(sigma, step, num_iter) + 1
    for m in range(len(size)):
        if kernel_size[m] >= size[m]:
            raise RuntimeError('resulting rectangular kernel size should be smaller w.r.t. specified grid: '
                               + str(kernel_size) + ' vs. ' + str(size))


# ---------------------------------------------------------------------------------------------

def barnes(pts, val, sigma, x0, step, size, method='optimized_convolution',
           num_iter=4, max_dist=3.5, min_weight=0.001, *, return_float64=False, exact=None, precision='fp64'):
    """
    Barnes interpolation of the observation values `val` at the sample points `pts` with
    Gaussian width `sigma` on the regular grid (`x0`, `step`, `size`) in 1, 2 or 3 dimensions.
    Signature, defaults, validation and result as `fastbarnes.interpolation.barnes`
    (reference interpolation.py:31-199): a new float32 array of shape `size[::-1]`
    (index order [y, x] / [z, y, x]), NaN where the nearest sample is farther than
    `max_dist * sigma`; the caller's arrays are never modified.

    Methods: 'optimized_convolution' (default) and 'convolution' (float32 result), and the
    exact O(N*W*H) Gaussian sums 'naive' and 'radius' (float64 result like the reference; every
    grid point sums the samples in sample order, 'radius' by exhaustive search instead of the
    reference's kd-tree, so they agree with the reference to rounding, not bit for bit).

    return_float64=True returns `(field32, field64)` where field64 is the fp64 quotient
    `vg/wg + offset` before the float32 cast (interpolation.py:367).

    exact (1D only): a 1D grid is a single line whose accumulator chains cannot be cut into pieces
    bit-exactly.  The default (None or True) walks the line with the 2 x num_iter (field, pass) chains
    side by side: bit-identical to the reference at any length (2^26 points in ~0.6 s).  exact=False
    opts into overlapping segments swept in parallel: ~10x faster for very long lines, but the result
    differs from the reference by the reference's own accumulated rounding (fp64 quotient within
    1e-8 x value range at 2^22 points; the NaN mask can flip where the weight sits on the threshold).
    """
    pts = _check_samples(pts, val)
    dim = pts.shape[1]
    sigma = _per_axis('sigma', sigma, dim)
    x0 = _per_axis('x0', x0, dim)
    step = _per_axis('step', step, dim)
    size = _grid_size(size, dim)
    max_dist_weight = exp(-max_dist ** 2 / 2)

    if method in _CONV_METHODS:
        _check_kernel_vs_grid(method, sigma, step, size, num_iter)
        flags = _precision_flag(precision, dim, return_float64)
        if dim == 1 and exact is False:
            flags |= FLAG_SEGMENTED_1D
        return _run(pts, val, sigma, x0, step, size, _CONV_METHODS[method], num_iter, max_dist_weight,
                    return_float64=return_float64, flags=flags)
    if method == 'radius':
        # specific checks of the reference (interpolation.py:187-193)
        if dim != 2:
            raise RuntimeError('radius algorithm works only in 2D but data is: ' + str(dim) + 'D')
        if sigma[0] != sigma[1]:
            raise RuntimeError('radius algorithm in 2D works only for scalar sigma value but sigma is: ' + str(sigma))
        return _run_exact(pts, val, sigma, x0, step, size, _lib.METHOD_RADIUS, max_dist_weight, min_weight)
    if method == 'naive':
        return _run_exact(pts, val, sigma, x0, step, size, _lib.METHOD_NAIVE, max_dist_weight, min_weight)
    raise RuntimeError("encountered invalid Barnes interpolation method: " + method)


def _run_exact(pts, val, sigma, x0, step, size, method_id, max_dist_weight, min_weight):
    """ The exact Gaussian sums 'naive' / 'radius' / 'naive_S2' (reference interpolation.py:862-938,
    :809-855, interpolationS2.py:260-301): float64 field of shape size[::-1]. """
    if pts.shape[0] == 0:
        raise ValueError('zero-size array to reduction operation minimum which has no identity')
    pts_c = np.ascontiguousarray(pts, dtype=np.float64)
    val_c = np.ascontiguousarray(val, dtype=np.float64)
    prob = _problem(len(size), sigma, x0, step, size, method_id, 1, max_dist_weight, 1, 0)
    out = np.empty(tuple(size[::-1]), dtype=np.float64)
    rc = _lib.lib().fb_barnes_exact_host(prob, pts_c.shape[0], _lib.dptr(pts_c), _lib.dptr(val_c),
                                         float(min_weight), _lib.dptr(out))
    _lib.check(rc)
    return out


def _run(pts, val, sigma, x0, step, size, method_id, num_iter, max_dist_weight, offsets=None, nfields=1,
         return_float64=False, flags=0):
    dim = len(size)
    if pts.shape[0] == 0:
        # np.amin of an empty array (reference: _normalize_values, :209)
        raise ValueError('zero-size array to reduction operation minimum which has no identity')
    pts_c = np.ascontiguousarray(pts, dtype=np.float64)
    val_c = np.ascontiguousarray(val, dtype=np.float64)
    prob = _problem(dim, sigma, x0, step, size, method_id, num_iter, max_dist_weight, nfields, flags)
    shape = tuple(size[::-1]) if nfields == 1 and offsets is None else (nfields,) + tuple(size[::-1])
    out = np.empty(shape, dtype=np.float32)
    out64 = np.empty(shape, dtype=np.float64) if return_float64 else None
    off_p = None
    if offsets is not None:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        off_p = offsets.ctypes.data_as(_lib.c_i64_p)
    rc = _lib.lib().fb_barnes_host(prob, pts_c.shape[0], off_p, pts_c.ctypes.data, val_c.ctypes.data,
                                   out.ctypes.data, out64.ctypes.data if return_float64 else None)
    _lib.check(rc)
    return (out, out64) if return_float64 else out


def barnes_batched(pts, val, sigma, x0, step, size, sample_offsets=None, method='optimized_convolution',
                   num_iter=4, max_dist=3.5, *, return_float64=False, precision='fp64'):
    """
    Interpolates B independent fields (e.g. time steps or ensemble members) on the same grid
    in one call.  `pts` (sum N_b, M) and `val` (sum N_b,) hold the samples of all fields
    back to back; field b owns rows [sample_offsets[b], sample_offsets[b+1]).  Alternatively
    pass `pts` of shape (B, N, M) and `val` of shape (B, N) with sample_offsets=None.
    Returns a float32 array of shape (B,) + size[::-1]; field b equals
    `barnes(pts_b, val_b, ...)` bit for bit (also with precision='fp32', see `barnes`).
    """
    if method not in _CONV_METHODS:
        raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))
    if not isinstance(pts, np.ndarray) or not isinstance(val, np.ndarray):
        raise RuntimeError('specified pts / val is not a numpy ndarray')
    if sample_offsets is None:
        if pts.ndim == 2 and val.ndim == 2:          # (B, N) one-dimensional fields
            pts = pts[:, :, None]
        if pts.ndim != 3 or val.ndim != 2 or pts.shape[:2] != val.shape:
            raise RuntimeError('expected pts of shape (B, N, M) and val of shape (B, N)')
        nfields, n, dim = pts.shape
        sample_offsets = np.arange(nfields + 1, dtype=np.int64) * n
        pts = pts.reshape(nfields * n, dim)
        val = val.reshape(nfields * n)
    else:
        sample_offsets = np.asarray(sample_offsets, dtype=np.int64)
        pts = _check_samples(pts, val)
        nfields = len(sample_offsets) - 1
        if nfields < 1 or sample_offsets[0] != 0 or sample_offsets[-1] != pts.shape[0] \
                or np.any(np.diff(sample_offsets) < 0):
            raise RuntimeError('sample_offsets must rise from 0 to the number of samples')
    dim = pts.shape[1]
    if dim < 1 or dim > 3:
        raise RuntimeError('Barnes interpolation supports only sample points in dimensions 1, 2 or 3')
    if np.any(np.diff(sample_offsets) == 0):
        raise ValueError('zero-size array to reduction operation minimum which has no identity')
    sigma = _per_axis('sigma', sigma, dim)
    x0 = _per_axis('x0', x0, dim)
    step = _per_axis('step', step, dim)
    size = _grid_size(size, dim)
    _check_kernel_vs_grid(method, sigma, step, size, num_iter)
    return _run(pts, val, sigma, x0, step, size, _CONV_METHODS[method], num_iter, exp(-max_dist ** 2 / 2),
                offsets=sample_offsets, nfields=nfields, return_float64=return_float64,
                flags=_precision_flag(precision, dim, return_float64))


# ---------------------------------------------------------------------------------------------
# the njit drivers of the reference, as thin calls into the same C ABI

def _interpolate_opt_convol(pts, val, sigma, x0, step, size, num_iter, max_dist_weight):
    """ Reference interpolation.py:329-367 (val is NOT modified here, unlike the reference). """
    return _run(np.asarray(pts).reshape(len(val), -1), val, sigma, x0, step, tuple(size),
                _lib.METHOD_OPTIMIZED_CONVOLUTION, num_iter, max_dist_weight)


def _interpolate_convol(pts, val, sigma, x0, step, size, num_iter, max_dist_weight):
    """ Reference interpolation.py:575-612 (val is NOT modified here, unlike the reference). """
    return _run(np.asarray(pts).reshape(len(val), -1), val, sigma, x0, step, tuple(size),
                _lib.METHOD_CONVOLUTION, num_iter, max_dist_weight)


def _inject_data(pts, val, x0, step, size):
    """
    Reference interpolation.py:205-212 + :219-322: centres the values and injects them.
    Returns (vg, wg, offset) with vg, wg float64 arrays of shape size[::-1].
    """
    size = tuple(int(s) for s in size)
    dim = len(size)
    pts_c = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, dim))
    val_c = np.ascontiguousarray(val, dtype=np.float64)
    ones = np.ones(dim)
    prob = _problem(dim, ones, _per_axis('x0', x0, dim), _per_axis('step', step, dim), size,
                    _lib.METHOD_OPTIMIZED_CONVOLUTION, 1, 0.0)
    vg = np.empty(size[::-1], dtype=np.float64)
    wg = np.empty(size[::-1], dtype=np.float64)
    offset = np.empty(1, dtype=np.float64)
    rc = _lib.lib().fb_inject_host(prob, pts_c.shape[0], None, _lib.dptr(pts_c), _lib.dptr(val_c),
                                   _lib.dptr(vg), _lib.dptr(wg), _lib.dptr(offset))
    _lib.check(rc)
    return vg, wg, float(offset[0])


def _accumulate_lines(in_arr, h_arr, arr_len, rect_len, num_iter, alpha):
    arr_len = int(arr_len)
    line = np.ascontiguousarray(in_arr[:arr_len], dtype=np.float64).copy()
    rc = _lib.lib().fb_accumulate_lines_host(_lib.dptr(line), 1, arr_len, 1, int(rect_len), int(num_iter),
                                             float(alpha))
    _lib.check(rc)
    # the reference ping-pongs between the two buffers and returns the one written last
    res = h_arr if (num_iter % 2) else in_arr
    res[:arr_len] = line
    return res


def _accumulate_tail_array(in_arr, h_arr, arr_len, rect_len, num_iter, alpha):
    """ Reference interpolation.py:485-533: n-fold tailed box filter of one line. """
    return _accumulate_lines(in_arr, h_arr, arr_len, rect_len, num_iter, alpha)


def _accumulate_array(in_arr, h_arr, arr_len, rect_len, num_iter):
    """ Reference interpolation.py:729-772: n-fold box filter of one line (tail value 0). """
    return _accumulate_lines(in_arr, h_arr, arr_len, rect_len, num_iter, 0.0)


def _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight):
    dim = len(size)
    ks = np.ascontiguousarray(kernel_size, dtype=np.int32)
    tv = np.zeros(dim) if tail_value is None else np.ascontiguousarray(tail_value, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    L = _lib.lib()
    csf = L.fb_conv_scale_factor(dim, ks.ctypes.data_as(_lib.c_i32_p), _lib.dptr(tv), _lib.dptr(sigma),
                                 _lib.dptr(step), int(num_iter), float(max_dist_weight))
    if not (vg.flags.c_contiguous and wg.flags.c_contiguous and vg.dtype == np.float64 and wg.dtype == np.float64):
        raise RuntimeError('vg and wg must be C-contiguous float64 arrays')
    sz = np.asarray(size, dtype=np.int64)
    rc = L.fb_convolve_host(dim, _lib.dptr(vg), _lib.dptr(wg), sz.ctypes.data_as(_lib.c_i64_p),
                            ks.ctypes.data_as(_lib.c_i32_p), int(num_iter), _lib.dptr(tv), csf)
    _lib.check(rc)


def _convolve_tail_1d(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight):
    """ Reference interpolation.py:373-394, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight)


def _convolve_tail_2d(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight):
    """ Reference interpolation.py:398-430, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight)


def _convolve_tail_3d(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight):
    """ Reference interpolation.py:434-479, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, tail_value, max_dist_weight)


def _convolve_1d(vg, wg, sigma, step, size, kernel_size, num_iter, max_dist_weight):
    """ Reference interpolation.py:617-639, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, None, max_dist_weight)


def _convolve_2d(vg, wg, sigma, step, size, kernel_size, num_iter, max_dist_weight):
    """ Reference interpolation.py:642-675, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, None, max_dist_weight)


def _convolve_3d(vg, wg, sigma, step, size, kernel_size, num_iter, max_dist_weight):
    """ Reference interpolation.py:678-724, in place on vg, wg. """
    _convolve(vg, wg, sigma, step, size, kernel_size, num_iter, None, max_dist_weight)


# ---------------------------------------------------------------------------------------------
# kernel parameters (reference interpolation.py:538-569, :777-803)

def _to_np(value):
    """ scalar -> one-element float64 array (:538-540). """
    return np.asarray([value], dtype=np.float64)


def _get_half_kernel_size_opt(sigma, step, num_iter):
    """ Half kernel size T of the tailed rectangular kernel, array version (:549-552). """
    L = _lib.lib()
    return np.asarray([L.fb_half_kernel_size_opt(float(s), float(d), int(num_iter))
                       for s, d in zip(np.atleast_1d(sigma), np.atleast_1d(step))], dtype=np.int32)


def _get_tail_value(sigma, step, num_iter):
    """ Tail value alpha, array version (:561-569). """
    L = _lib.lib()
    return np.asarray([L.fb_tail_value(float(s), float(d), int(num_iter))
                       for s, d in zip(np.atleast_1d(sigma), np.atleast_1d(step))], dtype=np.float64)


def _get_half_kernel_size(sigma, step, num_iter):
    """ Half kernel size T of the plain rectangular kernel, array version (:783-785). """
    L = _lib.lib()
    return np.asarray([L.fb_half_kernel_size(float(s), float(d), int(num_iter))
                       for s, d in zip(np.atleast_1d(sigma), np.atleast_1d(step))], dtype=np.int32)


def _get_sigma_effective(sigma, step, num_iter):
    """ Effective sigma of the n-fold plain rectangular kernel, array version (:797-803). """
    hks = _get_half_kernel_size(sigma, step, num_iter)
    return np.sqrt(num_iter / 3.0 * hks * (hks + 1)) * np.atleast_1d(np.asarray(step, dtype=np.float64))


def get_half_kernel_size_opt(sigma, step, num_iter):
    """ Scalar version (:543-545). """
    return _get_half_kernel_size_opt(_to_np(sigma), _to_np(step), num_iter)[0]


def get_tail_value(sigma, step, num_iter):
    """ Scalar version (:555-557). """
    return _get_tail_value(_to_np(sigma), _to_np(step), num_iter)[0]


def get_half_kernel_size(sigma, step, num_iter):
    """ Scalar version (:777-779). """
    return _get_half_kernel_size(_to_np(sigma), _to_np(step), num_iter)[0]


def get_sigma_effective(sigma, step, num_iter):
    """ Scalar version (:788-793). """
    return _get_sigma_effective(_to_np(sigma), _to_np(step), num_iter)[0]


# ---------------------------------------------------------------------------------------------
# device-resident interface (torch supplies device memory and streams; plumbing only)

def _check_sample_offsets(sample_offsets, nfields, nsamples):
    """ None, or nfields + 1 non-decreasing offsets from 0 to nsamples as a contiguous int64 array (the C side reads exactly
    nfields + 1 entries). """
    if sample_offsets is None:
        return None
    off = np.ascontiguousarray(sample_offsets, dtype=np.int64).reshape(-1)
    if len(off) != nfields + 1:
        raise RuntimeError('sample_offsets must have nfields + 1 = %d entries: %d' % (nfields + 1, len(off)))
    if off[0] != 0 or off[-1] != nsamples or np.any(np.diff(off) < 0):
        raise RuntimeError('sample_offsets must be non-decreasing, start at 0 and end at nsamples')
    return off


class BarnesDevice:
    """
    Device-resident plan for repeated interpolation of `nfields` fields with `nsamples_total`
    samples on a fixed grid: inputs and outputs are torch CUDA tensors, nothing is copied
    through the host, and calls only enqueue kernels on the current torch stream.

        plan = BarnesDevice(dim=2, sigma=1.0, x0=[...], step=1/32, size=(2400, 1200),
                            nfields=64, nsamples=64*50000, num_iter=4)
        out = plan(pts_dev, val_dev)            # float32 (64, 1200, 2400), stays on the GPU
    """

    def __init__(self, dim, sigma, x0, step, size, nfields, nsamples, method='optimized_convolution',
                 num_iter=4, max_dist=3.5, sample_offsets=None, device=None, want_float64=False, precision='fp64'):
        import torch
        if method not in _CONV_METHODS:
            raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device available; this package has no CPU fallback')
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.dim = dim
        sigma = _per_axis('sigma', sigma, dim)
        x0 = _per_axis('x0', x0, dim)
        step = _per_axis('step', step, dim)
        self.size = _grid_size(size, dim)
        _check_kernel_vs_grid(method, sigma, step, self.size, num_iter)
        self.nfields = int(nfields)
        self.nsamples = int(nsamples)
        self.prob = _problem(dim, sigma, x0, step, self.size, _CONV_METHODS[method], num_iter,
                             exp(-max_dist ** 2 / 2), nfields, _precision_flag(precision, dim, want_float64))
        self.offsets = _check_sample_offsets(sample_offsets, self.nfields, self.nsamples)
        L = _lib.lib()
        nbytes = L.fb_workspace_bytes(self.prob, self.nsamples)
        if nbytes < 0:
            _lib.check(int(nbytes))
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            shape = (self.nfields,) + tuple(self.size[::-1])
            self.out = torch.empty(shape, dtype=torch.float32, device=self.device)
            self.out64 = torch.empty(shape, dtype=torch.float64, device=self.device) if want_float64 else None
        self.workspace_bytes = int(nbytes)

    def __call__(self, pts, val, out=None):
        """ pts (nsamples, dim) float64 CUDA tensor, val (nsamples,) float64 CUDA tensor. """
        torch = self.torch
        if pts.dtype != torch.float64 or val.dtype != torch.float64 or not pts.is_cuda or not val.is_cuda:
            raise RuntimeError('pts and val must be float64 CUDA tensors')
        if not (pts.is_contiguous() and val.is_contiguous()):
            raise RuntimeError('pts and val must be contiguous')
        if pts.numel() != self.nsamples * self.dim or val.numel() != self.nsamples:
            raise RuntimeError('unexpected number of samples')
        if pts.device != self.device or val.device != self.device:
            raise RuntimeError('pts and val must live on the device of the plan (%s)' % (self.device,))
        if out is not None:
            shape = (self.nfields,) + tuple(self.size[::-1])
            if (not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != shape
                    or out.device != self.device):
                raise RuntimeError('out must be a contiguous float32 CUDA tensor of shape %s on %s' % (shape, self.device))
        out = self.out if out is None else out
        L = _lib.lib()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            off_p = None if self.offsets is None else self.offsets.ctypes.data_as(_lib.c_i64_p)
            rc = L.fb_barnes_dev(self.prob, self.nsamples, off_p, pts.data_ptr(), val.data_ptr(), out.data_ptr(),
                                 None if self.out64 is None else self.out64.data_ptr(),
                                 self.workspace.data_ptr(), self.workspace_bytes, stream)
        _lib.check(rc)
        return out
