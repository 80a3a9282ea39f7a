#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
q_lab.py -- GPU lab for the second-generation sweep kernels (csrc/fb_sweepq.cuh).

  python tools/q_lab.py check      bit-for-bit comparison q kernels vs first-generation kernels (and the oracle on a
                                   few cases) over a matrix of kernel widths, pass counts, shapes, 2D and 3D
  python tools/q_lab.py time       stage times of the bench workload (64 fields x 2400x1200, N=50000) under a list of
                                   option settings; one JSON line per setting

Results go to stdout and to gpurun_out/q_lab_<mode>.json.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    u = np.uint32 if a.dtype == np.float32 else np.uint64
    return bool(np.array_equal(a.view(u), b.view(u)))


def set_opts(L, **kw):
    from fastbarnes import _lib
    for k, v in kw.items():
        _lib.check(L.fb_set_option(k.encode(), int(v)))


def check():
    import torch
    from fastbarnes import interpolation as fb
    from fastbarnes import _lib
    from oracle import oracle as orc
    L = _lib.lib()
    rng = np.random.default_rng(4242)
    results = []
    ok_all = True
    # (dim, size, sigma / step per axis, nfields, num_iter list)
    cases = [
        (2, (512, 500), (8.0, 8.0), 6, (1, 2, 3, 4, 5, 6)),        # T=6..13: D%8 != 0 (mirrors), all-TMEM rings
        (2, (2400, 1200), (32.0, 32.0), 2, (4,)),                  # the paper kernel: T=27, D=56, one ring in shared memory
        (2, (333, 217), (13.0, 9.0), 3, (3, 4)),                   # ragged groups (217 % 16, 333 % 16 != 0), anisotropic
        (2, (1000, 300), (60.0, 5.0), 2, (4,)),                    # T=51 on x (4-warp plan), T=3 (D=8, the minimum) on y
        (2, (640, 400), (40.0, 34.0), 2, (2, 4, 6)),               # T around 30..40
        (3, (160, 128, 96), (6.0, 7.0, 6.5), 1, (1, 2, 4, 5)),     # 3D: transposing x, in-place y, finalising z
        (3, (70, 45, 50), (6.0, 5.0, 5.5), 2, (3,)),
    ]
    for dim, size, ratio, nf, iters in cases:
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
            try:
                plan = fb.BarnesDevice(dim, sigma, [0.0] * dim, step, size, nfields=nf, nsamples=nf * N, num_iter=n,
                                       want_float64=True)
            except RuntimeError as e:
                results.append({'case': [dim, size, ratio, n], 'skipped': str(e)[:80]})
                continue
            t0 = time.perf_counter()
            set_opts(L, sweepq=1)
            a = plan(d_pts, d_val).cpu().numpy()
            a64 = plan.out64.cpu().numpy()
            set_opts(L, sweepq=0)
            b = plan(d_pts, d_val).cpu().numpy()
            b64 = plan.out64.cpu().numpy()
            set_opts(L, sweepq=1)
            same = bits_equal(a, b) and bits_equal(a64, b64)
            rec = {'case': [dim, list(size), list(ratio), nf, n], 'q_equals_gen1': same,
                   'nan_frac': float(np.isnan(a).mean()), 'sec': round(time.perf_counter() - t0, 3)}
            if not same:
                bad = np.argwhere(a64.view(np.uint64) != b64.view(np.uint64))
                rec['n_diff'] = int(len(bad))
                rec['first_diff'] = [int(x) for x in bad[0]] if len(bad) else None
                d = np.abs(a64 - b64)
                rec['max_abs_diff'] = float(np.nanmax(d)) if np.isfinite(d).any() else None
                rec['nan_mismatch'] = int((np.isnan(a64) != np.isnan(b64)).sum())
                ok_all = False
            results.append(rec)
            print(json.dumps(rec), flush=True)
    # a few direct oracle comparisons (q path on)
    for dim, size, ratio, n in [(2, (512, 500), (8.0, 8.0), 4), (3, (70, 45, 50), (6.0, 5.0, 5.5), 3)]:
        step = 0.1
        sigma = [r * step for r in ratio]
        ext = (np.asarray(size) - 1) * step
        pts = rng.uniform(0, 1, (3000, dim)) * ext
        val = rng.normal(5, 2, 3000)
        a = fb.barnes(pts, val, sigma, [0.0] * dim, step, size, num_iter=n)
        ref = orc.barnes(pts, val, sigma, [0.0] * dim, step, size, num_iter=n, nthreads=8)
        rec = {'oracle_case': [dim, list(size), n], 'equal': bits_equal(a, ref)}
        ok_all = ok_all and rec['equal']
        results.append(rec)
        print(json.dumps(rec), flush=True)
    print(json.dumps({'all_ok': ok_all}))
    return results, ok_all


def timing(settings=None):
    import torch
    from fastbarnes import interpolation as fbi
    from fastbarnes import _lib
    import bench
    L = _lib.lib()
    F = int(os.environ.get('QLAB_FIELDS', '64'))
    pts_h, val_h = bench.make_fields(0, F)
    d_pts = torch.from_numpy(pts_h.reshape(F * bench.N_PER_FIELD, 2)).cuda()
    d_val = torch.from_numpy(val_h.reshape(F * bench.N_PER_FIELD)).cuda()
    plan = fbi.BarnesDevice(2, bench.SIGMA, bench.X0, bench.STEP, bench.SIZE, nfields=F, nsamples=F * bench.N_PER_FIELD,
                            num_iter=bench.NUM_ITER)
    if settings is None and os.environ.get('QLAB_SETTINGS'):
        settings = json.loads(os.environ['QLAB_SETTINGS'])
    if settings is None:
        settings = [
            dict(sweepq=0),
            dict(sweepq=1, sparse_inject=0),
            dict(sweepq=1, sparse_inject=1),
        ]
    ref = None
    results = []
    seg = np.zeros(5)
    nl = np.zeros(1, dtype=np.int64)
    for st in settings:
        set_opts(L, **st)
        for _ in range(3):
            out = plan(d_pts, d_val)
        torch.cuda.synchronize()
        L.fb_set_profiling(1)
        acc = np.zeros(5)
        steps = int(os.environ.get('QLAB_STEPS', '10'))
        for _ in range(steps):
            out = plan(d_pts, d_val)
            _lib.check(L.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
            acc += seg
        L.fb_set_profiling(0)
        acc /= steps
        o = out.cpu().numpy()
        refp = os.environ.get('QLAB_REF_OUT')
        if refp and ref is None:
            if os.path.exists(refp):
                ref = np.load(refp)                  # the field of the reference build (first setting: gen-1 kernels)
            else:
                np.save(refp, o)
        if ref is None:
            ref = o
        rec = dict(st)
        rec.update({'ms_zero': round(acc[0], 4), 'ms_inject': round(acc[1], 4), 'ms_x': round(acc[2], 4), 'ms_y': round(acc[3], 4),
                    'equal_to_first': bits_equal(o, ref)})
        results.append(rec)
        print(json.dumps(rec), flush=True)
    set_opts(L, sweepq=1, sparse_inject=1)
    return results


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'check'
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if mode == 'check':
        res, ok = check()
    else:
        res, ok = timing(), True
    with open(os.path.join(ROOT, 'gpurun_out', 'q_lab_%s.json' % mode), 'w') as f:
        json.dump(res, f, indent=1)
    sys.exit(0 if ok else 1)
