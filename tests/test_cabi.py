# -*- coding: utf-8 -*-
"""
CPU tests (no GPU): the C-ABI library loads and exports every symbol include/fastbarnes_b200.h
declares; the host-side arithmetic (T, alpha, conv_scale_factor, Lambert constants) is
bit-identical to the reference; the drop-in Python layer validates arguments like the reference
and fails loudly (RuntimeError) when no CUDA device is present -- there is no CPU fallback.
No compute entry point is exercised successfully here.
"""
import ctypes
import os
import re
from math import exp

import numpy as np
import pytest

from conftest import ROOT, load_golden
from fastbarnes import _lib, interpolation, interpolationS2
from fastbarnes.util import lambert_conformal


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'fastbarnes_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(fb_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), 'missing export: ' + s
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_problem_struct_layout():
    # fb_problem: 4 x int32, int64, 3 x int64, 9 x double, double
    assert ctypes.sizeof(_lib.FbProblem) == 16 + 8 + 24 + 72 + 8


def test_kernel_parameters_bit_identical_to_reference():
    L = _lib.lib()
    tab = load_golden('params')['table']
    mdw = exp(-3.5 ** 2 / 2)
    for sigma, step, n, T, Tp, alpha, csf, csfp in tab:
        n = int(n)
        assert L.fb_half_kernel_size_opt(sigma, step, n) == int(T)
        assert L.fb_half_kernel_size(sigma, step, n) == int(Tp)
        assert L.fb_tail_value(sigma, step, n) == alpha
        ks = np.asarray([2 * int(T) + 1], dtype=np.int32)
        got = L.fb_conv_scale_factor(1, ks.ctypes.data_as(_lib.c_i32_p), _lib.dptr(np.asarray([alpha])),
                                     _lib.dptr(np.asarray([sigma])), _lib.dptr(np.asarray([step])), n, mdw)
        assert got == csf
    # scalar / array helpers of the drop-in module (reference tests/BasicTest.py:112,132 use them)
    assert interpolation.get_half_kernel_size_opt(1.0, 1 / 32, 4) == 27
    assert interpolation.get_tail_value(1.0, 1 / 32, 4) == 0.20833333333333334
    assert interpolation.get_half_kernel_size(0.5, 1 / 32, 4) == int(np.sqrt(3.0 / 4) * 0.5 * 32 + 0.5)
    assert list(interpolation._get_half_kernel_size_opt(np.asarray([0.5, 1.0]), np.asarray([0.05, 0.125]), 3)) == [9, 7]
    assert interpolation.get_sigma_effective(1.0, 1 / 512, 4) > 0


def test_lambert_projection_constants():
    g = load_golden('s2_res8')
    assert np.array_equal(np.asarray(interpolationS2.get_lambert_proj()), g['proj'])
    assert lambert_conformal.create_proj(11.5, 34.5, 42.5, 65.5)[1] == 0.8146543793023436


def test_argument_validation_like_reference():
    pts = np.zeros((3, 2)); val = np.zeros(3)
    with pytest.raises(RuntimeError, match='not a numpy ndarray'):
        interpolation.barnes([[0.0, 0.0]], np.zeros(1), 1.0, [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='not a numpy ndarray'):
        interpolation.barnes(pts, [0.0, 0.0, 0.0], 1.0, [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='expected pts array'):
        interpolation.barnes(np.zeros((2, 2, 2)), np.zeros(2), 1.0, [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='dimensions 1, 2 or 3'):
        interpolation.barnes(np.zeros((2, 4)), np.zeros(2), 1.0, [0] * 4, 0.1, (5,) * 4)
    with pytest.raises(RuntimeError, match='expected val array'):
        interpolation.barnes(pts, np.zeros((3, 1)), 1.0, [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='inconsistent shapes'):
        interpolation.barnes(pts, np.zeros(4), 1.0, [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='sigma with invalid length'):
        interpolation.barnes(pts, val, [1.0, 1.0, 1.0], [0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='x0 with invalid length'):
        interpolation.barnes(pts, val, 1.0, [0, 0, 0], 0.1, (50, 50))
    with pytest.raises(RuntimeError, match='step with invalid length'):
        interpolation.barnes(pts, val, 1.0, [0, 0], [0.1], (50, 50))
    with pytest.raises(RuntimeError, match='size with invalid length'):
        interpolation.barnes(pts, val, 1.0, [0, 0], 0.1, (50, 50, 50))
    with pytest.raises(RuntimeError, match='array size should be array-like'):
        interpolation.barnes(pts, val, 1.0, [0, 0], 0.1, 50)
    with pytest.raises(RuntimeError, match='rectangular kernel size should be smaller'):
        interpolation.barnes(pts, val, 1.0, [0, 0], 0.1, (50, 17))       # sigma/step = 10, n = 4: T = 8, kernel 17
    with pytest.raises(RuntimeError, match='invalid Barnes interpolation method'):
        interpolation.barnes(pts, val, 1.0, [0, 0], 0.1, (50, 50), method='nope')
    # method-specific checks of 'radius' (reference interpolation.py:187-193)
    with pytest.raises(RuntimeError, match='radius algorithm works only in 2D'):
        interpolation.barnes(np.zeros((len(val), 3)), val, 1.0, [0, 0, 0], 0.1, (50, 50, 50), method='radius')
    with pytest.raises(RuntimeError, match='works only for scalar sigma'):
        interpolation.barnes(pts, val, [1.0, 2.0], [0, 0], 0.1, (50, 50), method='radius')
    with pytest.raises(RuntimeError, match='invalid Barnes interpolation method'):
        interpolationS2.barnes_S2(pts, val, 1.0, [0, 0], 0.1, (50, 50), method='nope')


def test_kernel_check_threshold_matches_reference():
    # reference tests/BasicTest.py:105-148: size == kernel_size must raise, kernel_size + 1 must pass validation
    sigma, step = 0.5, 1.0 / 32
    pts = np.asarray([0.2, 0.4, 0.7]); val = np.asarray([18.5, 17.0, 19.25])
    for method, hk in (('convolution', interpolation.get_half_kernel_size),
                       ('optimized_convolution', interpolation.get_half_kernel_size_opt)):
        for n in (3, 4, 6):
            ks = 2 * hk(sigma, step, n) + 1
            with pytest.raises(RuntimeError, match='rectangular kernel size'):
                interpolation.barnes(pts, val, sigma, 0.0, step, int(ks), method=method, num_iter=n)


@pytest.mark.skipif(_lib.lib().fb_device_count() > 0, reason='a CUDA device is present')
def test_no_cpu_fallback_without_gpu():
    pts = np.random.default_rng(0).uniform(0, 4, (10, 2)); val = np.ones(10)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        interpolation.barnes(pts, val, 0.5, [0.0, 0.0], 0.1, (64, 64))
    with pytest.raises(RuntimeError, match='no CUDA device'):
        interpolationS2.barnes_S2(pts, val, 1.0, [0.0, 0.0], 0.5, (20, 20), method='optimized_convolution_S2')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        interpolation._accumulate_tail_array(np.zeros(32), np.empty(32), 32, 7, 2, 0.5)
    assert _lib.lib().fb_set_device(0) == _lib.FB_ECUDA


def test_product_never_imports_the_oracle():
    """ the product path must not route through oracle/ """
    pkg = os.path.join(ROOT, 'fast-barnes-py_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')) or f == 'Makefile':
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.lower() or f == '_none_', os.path.join(dirpath, f)


def test_sample_offsets_are_validated_on_the_host():
    """ BarnesDevice hands sample_offsets to C code that reads exactly nfields + 1 entries: wrong lengths, offsets that do not
    run from 0 to nsamples or that decrease are refused before anything reaches the device (advisor finding, round 1). """
    from fastbarnes.interpolation import _check_sample_offsets
    assert _check_sample_offsets(None, 3, 10) is None
    ok = _check_sample_offsets([0, 4, 4, 10], 3, 10)
    assert ok.dtype == np.int64 and ok.flags['C_CONTIGUOUS'] and list(ok) == [0, 4, 4, 10]
    for bad in ([0, 4, 10], [0, 4, 4, 10, 10], [1, 4, 4, 10], [0, 4, 4, 9], [0, 5, 4, 10]):
        with pytest.raises(RuntimeError, match='sample_offsets'):
            _check_sample_offsets(bad, 3, 10)
