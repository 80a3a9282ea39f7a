# -*- coding: utf-8 -*-
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'fast-barnes-py_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


def bits_equal(a, b):
    """ Bit-for-bit equality of two float arrays (NaN payloads are treated as equal NaNs). """
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    ia = a.view(np.uint64 if a.dtype == np.float64 else np.uint32)
    ib = b.view(ia.dtype)
    same = (ia == ib) | (np.isnan(a) & np.isnan(b))
    return bool(np.all(same))


def same_up_to_zero_sign(a, b):
    """ Bit equality, except that +0.0 and -0.0 compare equal. """
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope='session')
def golden():
    return load_golden


CASES = ['case_1d_n4', 'case_1d_n6_plain', 'case_2d_n4', 'case_2d_aniso_n3', 'case_2d_n5_plain', 'case_2d_T0',
         'case_3d_n4', 'case_3d_n2']
