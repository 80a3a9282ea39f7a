# -*- coding: utf-8 -*-
"""
ctypes binding of the C ABI declared in include/fastbarnes_b200.h
(csrc/_build/libfastbarnes_b200.so, hand-written sm_100a CUDA kernels).

There is no CPU fallback: if the shared library cannot be found (and cannot be built
because nvcc is absent) importing this module fails, and every compute call raises
RuntimeError when no CUDA device is present.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_HERE, '..', 'csrc'))
LIB_PATH = os.environ.get('FB_LIB_PATH') or os.path.join(CSRC, '_build', 'libfastbarnes_b200.so')

FB_OK, FB_EINVAL, FB_ECUDA, FB_ENOMEM, FB_EKERNEL = 0, -1, -2, -3, -4
METHOD_OPTIMIZED_CONVOLUTION, METHOD_CONVOLUTION = 0, 1
METHOD_NAIVE, METHOD_RADIUS, METHOD_NAIVE_S2 = 2, 3, 4

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_i32_p = ctypes.POINTER(ctypes.c_int32)


class FbProblem(ctypes.Structure):
    """ struct fb_problem (include/fastbarnes_b200.h). """
    _fields_ = [('dim', ctypes.c_int32), ('method', ctypes.c_int32), ('num_iter', ctypes.c_int32),
                ('flags', ctypes.c_int32), ('nfields', ctypes.c_int64), ('size', ctypes.c_int64 * 3),
                ('sigma', ctypes.c_double * 3), ('x0', ctypes.c_double * 3), ('step', ctypes.c_double * 3),
                ('max_dist_weight', ctypes.c_double)]


class FbS2Map(ctypes.Structure):
    """ struct fb_s2_map (include/fastbarnes_b200.h). """
    _fields_ = [('proj', ctypes.c_double * 5), ('lam_x0', ctypes.c_double * 2), ('lam_extent', ctypes.c_double * 2)]


# every symbol include/fastbarnes_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'fb_last_error': (ctypes.c_char_p, []),
    'fb_version': (ctypes.c_int, []),
    'fb_device_count': (ctypes.c_int, []),
    'fb_set_device': (ctypes.c_int, [ctypes.c_int]),
    'fb_half_kernel_size_opt': (ctypes.c_int32, [ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    'fb_half_kernel_size': (ctypes.c_int32, [ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    'fb_tail_value': (ctypes.c_double, [ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    'fb_conv_scale_factor': (ctypes.c_double, [ctypes.c_int, c_i32_p, c_double_p, c_double_p, c_double_p,
                                               ctypes.c_int, ctypes.c_double]),
    'fb_barnes_host': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, c_i64_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'fb_workspace_bytes': (ctypes.c_int64, [ctypes.POINTER(FbProblem), ctypes.c_int64]),
    'fb_barnes_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, c_i64_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_halo_planes': (ctypes.c_int64, [ctypes.POINTER(FbProblem)]),
    'fb_slab_interleaved': (ctypes.c_int, [ctypes.POINTER(FbProblem)]),
    'fb_slab_layout': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int, c_i64_p, c_i64_p, c_i64_p]),
    'fb_slab_phase1_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_inject_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_sweeps_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_phase2_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_phase2_inplace_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_void_p]),
    'fb_slab_result_offsets': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_int, c_i64_p, c_i64_p]),
    'fb_ipc_event_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    'fb_ipc_event_open': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'fb_event_record': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'fb_stream_wait_event': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'fb_event_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'fb_accumulate_lines_host': (ctypes.c_int, [c_double_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_int64, ctypes.c_int, ctypes.c_double]),
    'fb_convolve_host': (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_i64_p, c_i32_p, ctypes.c_int,
                                        c_double_p, ctypes.c_double]),
    'fb_inject_host': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, c_i64_p, c_double_p,
                                      c_double_p, c_double_p, c_double_p, c_double_p]),
    'fb_barnes_exact_host': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, c_double_p, c_double_p,
                                            ctypes.c_double, c_double_p]),
    'fb_barnes_exact_dev': (ctypes.c_int, [ctypes.POINTER(FbProblem), ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'fb_lambert_create_proj': (ctypes.c_int, [ctypes.c_double] * 4 + [c_double_p]),
    'fb_lambert_to_map_host': (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int64, c_double_p]),
    'fb_s2_part1_host': (ctypes.c_int, [ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                        ctypes.c_int, ctypes.c_double, c_double_p, c_float_p]),
    'fb_s2_resample_host': (ctypes.c_int, [c_float_p, ctypes.c_int64, ctypes.c_int64, c_double_p, c_double_p,
                                           c_double_p, c_i64_p, c_double_p, c_float_p]),
    'fb_barnes_s2_host': (ctypes.c_int, [ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                         c_double_p, c_i64_p, ctypes.c_int, ctypes.c_double, c_double_p,
                                         c_float_p]),
    'fb_s2_default_map': (ctypes.c_int, [ctypes.POINTER(FbS2Map)]),
    'fb_s2_part1_map_host': (ctypes.c_int, [ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                            ctypes.c_int, ctypes.c_double, ctypes.POINTER(FbS2Map), c_float_p]),
    'fb_barnes_s2_map_host': (ctypes.c_int, [ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                             c_double_p, c_i64_p, ctypes.c_int, ctypes.c_double,
                                             ctypes.POINTER(FbS2Map), c_float_p]),
    'fb_set_option': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    'fb_kernel_launch_count': (ctypes.c_int64, []),
    'fb_set_profiling': (ctypes.c_int, [ctypes.c_int]),
    'fb_last_profile': (ctypes.c_int, [c_double_p, ctypes.c_int, c_i64_p]),
}


def _source_files():
    """ Everything the shared library is compiled from. """
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh')) or f == 'Makefile']
    files.append(os.path.normpath(os.path.join(CSRC, '..', '..', 'include', 'fastbarnes_b200.h')))
    return files


def _source_digest():
    import hashlib
    h = hashlib.sha256()
    for f in _source_files():
        h.update(os.path.basename(f).encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stamp_path():
    return os.path.join(os.path.dirname(LIB_PATH), 'sources.sha256')


def is_stale():
    """ True when the shared library is missing or was built from other sources (content hash, not mtimes:
    a copied tree keeps its binary, an edited kernel never runs a stale one). """
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(_stamp_path()) as f:
            return f.read().strip() != _source_digest()
    except OSError:
        return True


def mark_built():
    """ Records that LIB_PATH was just built from the current sources (after a manual `make`). """
    with open(_stamp_path(), 'w') as f:
        f.write(_source_digest() + '\n')


def build(force=False):
    """ Compiles csrc/ for sm_100a with nvcc (in-tree, csrc/_build/) when the sources changed. """
    if force or is_stale():
        nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
        if not os.path.exists(nvcc):
            raise RuntimeError('libfastbarnes_b200.so is missing or stale and nvcc was not found to build it')
        subprocess.check_call(['make', '-C', CSRC, '-s', '-B', 'NVCC=' + nvcc])
        mark_built()
    return LIB_PATH


_lib = None


def lib():
    """ Returns the loaded shared library with argument types set (loads / builds on first use). """
    global _lib
    if _lib is None:
        if os.environ.get('FB_LIB_PATH') is None:
            build()                                  # no-op unless the sources changed since the last build
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError('cannot load the CUDA library %s: %s (no CPU fallback exists)' % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        # tuning switches from the environment (results are bit-identical either way)
        if os.environ.get('FB_HOST_CHUNK_FIELDS') is not None:
            L.fb_set_option(b'host_chunk_fields', int(os.environ['FB_HOST_CHUNK_FIELDS']))
        for env, opt in (('FB_SWEEPQ', b'sweepq'), ('FB_SWEEPQ_STAGES', b'sweepq_stages'),
                         ('FB_SWEEPQ_PREFETCH', b'sweepq_prefetch'), ('FB_SWEEPQ_WARPS', b'sweepq_warps')):
            if os.environ.get(env) is not None:
                L.fb_set_option(opt, int(os.environ[env]))
        # any option: FB_OPTIONS="name=value,name=value"
        for item in os.environ.get('FB_OPTIONS', '').split(','):
            if '=' in item:
                k, v = item.split('=', 1)
                check(L.fb_set_option(k.strip().encode(), int(v)))
        _lib = L
    return _lib


def last_error():
    msg = lib().fb_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(rc):
    """ Maps a C return code to the RuntimeError the reference would raise. """
    if rc != FB_OK:
        raise RuntimeError(last_error() or ('fastbarnes_b200 error code %d' % rc))


def dptr(a):
    return a.ctypes.data_as(c_double_p)
