#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py -- headline benchmark: grid-points/s of the 2D optimized-convolution Barnes
interpolation (n=4) on the paper grid, batched independent fields (BASELINE.json configs[4]
shape: 2400x1200, step 1/32, sigma 1.0, N=50 000 samples per field).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--fields F] [--impl reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N>1).  A step = one pass of
the hot path (centre -> inject -> x sweep -> y sweep + mask/divide/cast) over a batch of F fields
per GPU (default 1024), run as sub-batches of 64 fields (one workspace of 5.9 GB fp64 state each)
that rotate over four distinct sample sets.  `value` is measured with the samples resident in HBM (CUDA events, max over ranks);
`e2e` goes through the HOST-buffer C-ABI call (pinned host memory; H2D of the samples and D2H
of the float32 fields inside the timed region).  Per-kernel times for the roofline come from
CUDA events recorded by the library on the launching stream during the same timed steps.

`--impl reference` times the CPU implementation of the same path (the oracle port of the
reference's Numba code, one field per host thread) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))

METRIC = 'grid-points/s for 2D optimized_convolution n=4'
UNIT = 'grid-points/s'
SIZE = (2400, 1200)
STEP = 1.0 / 32
X0 = np.asarray([-26.0 + STEP, 34.5])
SIGMA = 1.0
NUM_ITER = 4
N_PER_FIELD = 50000
POINTS_PER_FIELD = SIZE[0] * SIZE[1]
# algorithmic bytes per grid point, fp64 state, two fields (SURVEY.md section 8d / BASELINE.md section 4)
BYTES_ZERO = 16      # zero-fill of vg, wg
BYTES_SWEEP_X = 32   # non-final fused sweep: 2 x 8 B read + 2 x 8 B write
BYTES_SWEEP_Y = 20   # final fused sweep: 2 x 8 B read + 4 B float32 write
BYTES_TOTAL_2D = 68
SUB_FIELDS = 64      # fields per sub-batch (one call of the device-resident plan / of fb_barnes_host)
N_SETS = 4           # distinct sample sets the sub-batches rotate over


_JSON_OUT = None   # set by main(): duplicate of the original stdout

def make_fields(first_field, nfields):
    """ synthetic samples of fields [first_field, first_field+nfields): SURVEY.md section 8d, config C5 """
    pts = np.empty((nfields, N_PER_FIELD, 2))
    val = np.empty((nfields, N_PER_FIELD))
    for i in range(nfields):
        rng = np.random.default_rng(2000 + first_field + i)
        pts[i] = X0 + rng.uniform(0, 1, (N_PER_FIELD, 2)) * np.asarray([(SIZE[0] - 1) / 32, (SIZE[1] - 1) / 32])
        val[i] = rng.normal(1000, 10, N_PER_FIELD)
    return pts, val


def config_dict(fields_per_gpu, n_gpus):
    sub = min(SUB_FIELDS, fields_per_gpu)
    return {'workload': '2D optimized_convolution, %dx%d grid, step 1/32, sigma 1.0, num_iter 4, batched fields, '
                        'N=%d samples/field (BASELINE configs[4] shape; configs[0] grid)' % (SIZE[0], SIZE[1], N_PER_FIELD),
            'fields_per_gpu_per_step': fields_per_gpu, 'fields_per_step': fields_per_gpu * n_gpus,
            'grid': list(SIZE), 'samples_per_field': N_PER_FIELD, 'parallelism': 'fields sharded, no collective',
            'sub_batch_fields': sub, 'sample_sets': N_SETS,
            'cache': 'working set per sub-batch (%.1f GB fp64 state) exceeds L2; no flush needed'
                     % (sub * POINTS_PER_FIELD * 32 / 1e9),
            'streams': 'device-resident arm: consecutive sub-batches alternate over 2 CUDA streams (own workspace each)'}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, one field per thread

def cpu_fields_per_second(nfields, nthreads, repeats=1):
    from oracle import oracle as orc
    from concurrent.futures import ThreadPoolExecutor
    pts, val = make_fields(0, nfields)
    orc.lib()

    def one(i):
        return orc.barnes(pts[i], val[i], SIGMA, X0, STEP, SIZE, num_iter=NUM_ITER, nthreads=1)

    ref0 = one(0)   # warm-up (page faults, library load); also the parity reference of bench.py
    best = None
    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            list(ex.map(one, range(nfields)))
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
    return nfields / best, best, ref0


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nfields = max(cores, 8)            # bounded sample of the step's fields: one per host thread
    from oracle import oracle as orc
    from concurrent.futures import ThreadPoolExecutor
    pts, val = make_fields(0, nfields)
    orc.lib()

    def one(i):
        return orc.barnes(pts[i], val[i], SIGMA, X0, STEP, SIZE, num_iter=NUM_ITER, nthreads=1)

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(one, range(nfields)))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = nfields * POINTS_PER_FIELD * len(times) / total
    sample = ('%d fields per step (one per host thread) out of the %d fields of a step of the same 2400x1200 / N=50000 '
              'workload' % (nfields, args.fields * args.gpus))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args.fields, args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """ SM clock and throttle reasons of one GPU during the timed region: NVML polled every 10 ms from a thread
    (pynvml), nvidia-smi -lms as the fallback. """
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []          # (time, sm_mhz, set of reasons)
        self.proc = None
        self.nvml = None
        self.smmax = None
        self.gpu_index = gpu_index
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = ((n.nvmlClocksThrottleReasonHwSlowdown, 'hw_slowdown'),
                 (n.nvmlClocksThrottleReasonHwThermalSlowdown, 'hw_thermal_slowdown'),
                 (n.nvmlClocksThrottleReasonSwThermalSlowdown, 'sw_thermal_slowdown'),
                 (n.nvmlClocksThrottleReasonSwPowerCap, 'sw_power_cap'))
        while not self.stop_flag:
            try:
                clk = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), clk, {name for bit, name in names if mask & bit}))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(',')]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                self.smmax = float(f[2])
            except ValueError:
                continue
            reasons = {name for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9])
                       if v.lower().startswith('active')}
            self.rows.append((time.perf_counter(), clk, reasons))

    def stop(self, t_begin, t_end):
        if self.nvml is None and self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML / nvidia-smi']}
        time.sleep(0.05)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        inside = [(c, r) for (t, c, r) in self.rows if t_begin <= t <= t_end]
        if not inside:   # timed region shorter than the sampling period: fall back to all samples
            inside = [(c, r) for (t, c, r) in self.rows]
        reasons = set()
        for _, r in inside:
            reasons |= r
        sm = [c for c, _ in inside]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_min_mhz': float(min(sm)) if sm else None,
                'sm_max_mhz': self.smmax, 'reasons': sorted(reasons), 'samples': len(sm),
                'source': 'NVML polled every 10 ms' if self.nvml is not None else 'nvidia-smi -lms'}


# ------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, local_rank):
    """ Multi-GPU runs: pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the PCI device) before
    it allocates pinned host buffers, so that those land in the GPU's NUMA node and the end-to-end arm's
    host<->device copies of the ranks do not all cross the socket interconnect.  Returns the CPU list used. """
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bus) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception as e:                       # no sysfs / no permission: run unbound
        sys.stderr.write('[bench] NUMA binding skipped: %r\n' % (e,))
    return None


def extra_blocks(torch, fbi, L, dev):
    """ Secondary measurements on one GPU: the literal paper case (BASELINE configs[0]: ONE 2400x1200 field, N=3490,
    device-resident) and the S2 path at resolution 64 (BASELINE configs[3]: part 1 = projection + convolution on the
    4096x2816 Lambert grid, part 2 = resampling to 4800x2400; host API, copies included) """
    from math import exp
    from fastbarnes import interpolationS2 as fbs2
    out = {}
    try:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_paper.npz'))
        pts, val = np.ascontiguousarray(g['pts'], dtype=np.float64), np.ascontiguousarray(g['val'], dtype=np.float64)
    except Exception:
        rng = np.random.default_rng(7)
        pts = X0 + rng.uniform(0, 1, (3490, 2)) * np.asarray([(SIZE[0] - 1) / 32, (SIZE[1] - 1) / 32])
        val = rng.normal(1000, 10, 3490)
    n = len(val)
    plan = fbi.BarnesDevice(2, SIGMA, X0, STEP, SIZE, nfields=1, nsamples=n, num_iter=NUM_ITER, device=dev)
    d_p, d_v = torch.from_numpy(pts).to(dev), torch.from_numpy(val).to(dev)
    for _ in range(5):
        plan(d_p, d_v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        plan(d_p, d_v)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    out['c1_single_field'] = {'workload': 'ONE 2400x1200 field, N=%d, num_iter 4, device-resident (46 MB of fp64 state: L2 resident, '
                                          'launch / latency bound)' % n,
                              'us_per_field': us, 'value': POINTS_PER_FIELD / (us * 1e-6), 'unit': UNIT}
    # S2, resolution 64 (demo/timing5_table5.py:100-102, :131-152 times the two parts separately)
    step = 1.0 / 64
    x0 = np.asarray([-26.0 + step, 34.5])
    size = (4800, 2400)
    mdw = exp(-3.5 ** 2 / 2)
    sig = np.full(2, 1.0)
    stp = np.full(2, step)
    best1 = best2 = None
    for _ in range(3):
        t0 = time.perf_counter()
        res1 = fbs2.interpolate_opt_convol_S2_part1(pts, val, sig, x0, stp, size, NUM_ITER, mdw)
        t1 = time.perf_counter()
        fbs2.interpolate_opt_convol_S2_part2(*res1)
        t2 = time.perf_counter()
        best1 = t1 - t0 if best1 is None or t1 - t0 < best1 else best1
        best2 = t2 - t1 if best2 is None or t2 - t1 < best2 else best2
    out['c4_s2'] = {'workload': 'barnes_S2 optimized_convolution_S2, N=%d, 4800x2400 lon/lat grid at 1/64 degree (Lambert grid '
                                '4096x2816, T=54), host API with copies' % n,
                    'ms_part1_projection_and_convolution': best1 * 1e3, 'ms_part2_resampling': best2 * 1e3,
                    'reference_numba_seconds': {'part1': 1.200, 'part2': 0.381, 'source': 'SURVEY.md section 6.2 (build container, 1 thread)'}}
    return out


def slab3d_block(torch, dist, dev, rank, world, steps=5, warmup=2):
    """ Secondary measurement on ALL ranks: BASELINE configs[2] (C3) -- ONE 1024x1024x512 volume, N=1e7 samples, sigma 8 grid
    steps, num_iter 4 -- split into z-slabs over the ranks (fastbarnes.distributed.BarnesSlab3D: injection and x / y sweeps of
    the own planes, halo exchange over NCCL overlapped with the interior sweeps, z sweep + mask + divide + cast).  STRONG
    scaling: the volume is fixed, the time is the max over ranks of CUDA-event times of whole calls.  Every rank also runs
    the undivided volume once and compares its own planes (fp64 quotient; |diff| <= 1e-12 * value range, same NaN mask). """
    from fastbarnes import interpolation as fbi
    from fastbarnes import distributed as fd
    W, H, D, N, sigma, n_iter = 1024, 1024, 512, 10_000_000, 8.0, 4
    rng = np.random.default_rng(1235)                           # the same samples on every rank
    pts = rng.uniform(0.0, 1.0, (N, 3)) * np.asarray([W - 1, H - 1, D - 1], dtype=np.float64)
    val = rng.normal(0.0, 1.0, N)
    dp, dv = torch.from_numpy(pts).to(dev), torch.from_numpy(val).to(dev)
    slab = fd.BarnesSlab3D(sigma, [0.0] * 3, 1.0, (W, H, D), N, num_iter=n_iter, want_float64=True, device=dev,
                           exchange=os.environ.get('FB_SLAB_EXCHANGE', 'nccl'))
    exchange_mode = slab.exchange_mode

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        slab(dp, dv)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        slab(dp, dv)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    # the steps one by one (synchronised between them: no overlap) -- where the time goes
    parts = np.zeros(4)
    for _ in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        barrier()
        ev[0].record(); slab.inject(dp, dv)
        ev[1].record(); slab.sweeps(0, slab.zc)
        ev[2].record(); slab.exchange()
        ev[3].record(); slab.phase2()
        ev[4].record()
        barrier()
        parts += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    parts /= 3
    own64 = slab.out64.clone()
    own32 = slab.out.clone()
    halo, z0, z1 = slab.halo, slab.z0, slab.z1
    sent = sum((s_[1] - s_[0]) for _, s_, _ in slab.transfers() if s_) * W * H * 16
    del slab
    torch.cuda.empty_cache()
    # the undivided volume on this GPU
    plan = fbi.BarnesDevice(3, sigma, [0.0] * 3, 1.0, (W, H, D), nfields=1, nsamples=N, num_iter=n_iter, want_float64=True, device=dev)
    plan(dp, dv)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    plan(dp, dv)
    f1.record()
    torch.cuda.synchronize()
    ms_single = f0.elapsed_time(f1)
    ref64 = plan.out64.view(D, H, W)[z0:z1]
    ref32 = plan.out.view(D, H, W)[z0:z1]
    nan_same = bool(torch.equal(torch.isnan(own64), torch.isnan(ref64)))
    diff = float(torch.nan_to_num(own64 - ref64, nan=0.0).abs().max())
    vrange = float(dv.max() - dv.min())
    f32_same = float((own32.view(torch.int32) == ref32.view(torch.int32)).double().mean())
    t = torch.tensor([ms, diff, 0.0 if nan_same else 1.0, 1.0 - f32_same, ms_single] + list(parts), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, diff, nan_bad, f32_diff_frac, ms_single = (float(x) for x in t[:5])
    parts = [float(x) for x in t[5:]]
    del plan
    torch.cuda.empty_cache()
    return {'workload': 'ONE 1024x1024x512 volume, N=1e7 samples, sigma 8 grid steps, num_iter 4, fp64, device-resident samples; '
                        'z-slabs over %d GPU(s), halo %d planes per side' % (world, halo),
            'scaling': 'strong', 'n_gpus': world, 'halo_transport': exchange_mode, 'ms_per_volume': ms, 'value': W * H * D / (ms * 1e-3), 'unit': UNIT,
            'steps': steps, 'warmup': warmup,
            'ms_single_gpu_undivided': ms_single,
            'ms_steps_serialised_max_over_ranks': {'inject': parts[0], 'sweeps_xy': parts[1], 'halo_exchange': parts[2],
                                                   'sweep_z_finalise': parts[3]},
            'halo_bytes_sent_rank0': int(sent),
            'parity_checked': bool(nan_bad == 0.0 and diff <= 1e-12 * vrange),
            'parity': {'max_abs_diff_fp64_quotient': diff, 'bound': 1e-12 * vrange, 'nan_mask_identical': nan_bad == 0.0,
                       'float32_bits_differing_fraction': f32_diff_frac,
                       'against': 'the undivided volume on one GPU (bit-identical to the oracle, tests/test_gpu_parity.py)'}}


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from fastbarnes import interpolation as fbi
    from fastbarnes import _lib

    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().fb_set_device(local_rank))
    dev = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    F = args.fields
    L = _lib.lib()

    # ---- device-resident arm ---------------------------------------------------------------------
    SUB = min(SUB_FIELDS, F)
    nsub = (F + SUB - 1) // SUB                      # sub-batches per step
    F = nsub * SUB
    # N_SETS distinct sample sets of SUB fields each (field seeds 2000 + rank * N_SETS * SUB + ...); the sub-batches
    # of a step rotate over them, so consecutive launches never see the same samples
    sets_h = [make_fields((rank * N_SETS + k) * SUB, SUB) for k in range(N_SETS)]
    pin = [(torch.from_numpy(p_.reshape(SUB * N_PER_FIELD, 2)).pin_memory(),
            torch.from_numpy(v_.reshape(SUB * N_PER_FIELD)).pin_memory()) for (p_, v_) in sets_h]
    d_sets = [(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)) for (a, b) in pin]
    # one plan (workspace + output buffer) per stream: consecutive sub-batches alternate between the
    # streams so that the zero-fill / injection of sub-batch i+1 overlaps the sweeps of sub-batch i
    nstreams = max(1, args.streams)
    plans = [fbi.BarnesDevice(2, SIGMA, X0, STEP, SIZE, nfields=SUB, nsamples=SUB * N_PER_FIELD, num_iter=NUM_ITER, device=dev)
             for _ in range(nstreams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    plan = plans[0]
    torch.cuda.synchronize()

    def run_steps(n):
        """ n steps = n * nsub sub-batches, round-robin over the streams and the sample sets; returns after enqueueing """
        cur = torch.cuda.current_stream()
        for s_ in streams:
            s_.wait_stream(cur)
        for i in range(n * nsub):
            with torch.cuda.stream(streams[i % nstreams]):
                plans[i % nstreams](*d_sets[i % N_SETS])
        for s_ in streams:
            cur.wait_stream(s_)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(args.warmup)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    L.fb_set_profiling(1)
    seg = np.zeros(5)
    seg_ms = np.zeros(5)
    nl = np.zeros(1, dtype=np.int64)
    launches0 = L.fb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = L.fb_kernel_launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end)
    # per-kernel durations: the library's per-stage CUDA events (recorded on the launching stream) of single
    # sub-batches, read back after every call (reading forces a sync, so this is a separate loop from `value`)
    nprof = max(args.steps, 8)
    for i in range(nprof):
        plan(*d_sets[i % N_SETS])
        _lib.check(L.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
        seg_ms += seg
    seg_ms /= nprof
    L.fb_set_profiling(0)
    # field 0 of sample set 0 as the timed arm computes it (checked against the oracle below)
    check_dev = plan(*d_sets[0])[0].cpu().numpy() if rank == 0 else None

    # ---- secondary: the same batches in fp32 working precision (north_star's fp32 path; not the headline,
    # the reference computes in fp64) ----------------------------------------------------------------
    fp32 = None
    if not args.no_fp32:
        plans32 = [fbi.BarnesDevice(2, SIGMA, X0, STEP, SIZE, nfields=SUB, nsamples=SUB * N_PER_FIELD, num_iter=NUM_ITER,
                                    device=dev, precision='fp32') for _ in range(nstreams)]
        plans64, plans[:] = list(plans), plans32
        run_steps(min(args.warmup, 2))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(args.steps)
        e1.record()
        barrier()
        ms32 = e0.elapsed_time(e1)
        L.fb_set_profiling(1)
        seg32 = np.zeros(5)
        for i in range(nprof):
            plans32[0](*d_sets[i % N_SETS])
            _lib.check(L.fb_last_profile(seg.ctypes.data_as(_lib.c_double_p), 5, nl.ctypes.data_as(_lib.c_i64_p)))
            seg32 += seg
        seg32 /= nprof
        L.fb_set_profiling(0)
        plans[:] = plans64
        fp32 = (ms32, seg32)
        del plans32

    # ---- secondary blocks (rank 0, single GPU): the literal paper field and the S2 path -----------------
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = extra_blocks(torch, fbi, L, dev)

    # ---- secondary block on all ranks: the C3 volume as z-slabs (strong scaling) --------------------------
    slab3d = None
    if not args.no_extra:
        try:
            slab3d = slab3d_block(torch, dist, dev, rank, world)
        except Exception as e:                       # a secondary block must not take the headline line with it
            if world > 1:
                raise                                # ... but a rank that drops out of the collectives must not hang the others
            slab3d = {'error': repr(e)[:300]}

    # ---- end-to-end arm: host buffers through the C ABI -----------------------------------------------
    # One call of the public entry point takes EK sub-batches (up to 256 fields: the sample sets back to back), so that a
    # step is a handful of calls; every call returns with its results in host memory.
    ek = 4 if nsub % 4 == 0 else (2 if nsub % 2 == 0 else 1)
    ek = min(ek, N_SETS)
    EF = ek * SUB                                    # fields per call
    ncalls = nsub // ek
    e_pts = torch.cat([pin[k][0] for k in range(ek)]).pin_memory()
    e_val = torch.cat([pin[k][1] for k in range(ek)]).pin_memory()
    out_pin = torch.empty((EF,) + SIZE[::-1], dtype=torch.float32).pin_memory()
    prob = type(plan.prob).from_buffer_copy(plan.prob)
    prob.nfields = EF
    h2d = ncalls * (e_pts.numel() * 8 + e_val.numel() * 8)
    d2h = ncalls * out_pin.numel() * 4

    def e2e_step():
        for _ in range(ncalls):
            _lib.check(L.fb_barnes_host(prob, EF * N_PER_FIELD, None, e_pts.data_ptr(), e_val.data_ptr(), out_pin.data_ptr(), None))

    e2e_step()
    check_e2e = out_pin[0].numpy().copy() if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- what the host link allows: the result buffer of one sub-batch copied device -> pinned host, all ranks at
    # the same time (the end-to-end arm moves 4 B per grid point this way and almost nothing the other way) ------------
    d_out = plan(*d_sets[0])
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_reps = 4
    probe_dst = out_pin[:SUB]
    probe_dst.copy_(d_out, non_blocking=True)
    d2h_ms = None
    for _ in range(3):                               # the best of three rounds: a ceiling, not an average
        barrier()
        p0.record()
        for _ in range(probe_reps):
            probe_dst.copy_(d_out, non_blocking=True)
        p1.record()
        barrier()
        ms_round = p0.elapsed_time(p1) / probe_reps
        d2h_ms = ms_round if d2h_ms is None or ms_round < d2h_ms else d2h_ms

    # ---- reduce over ranks (max time) ------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_s * 1e3, fp32[0] if fp32 else 0.0, d2h_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms32_total, d2h_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank == 0:
        pts_per_step = F * POINTS_PER_FIELD * world
        value = pts_per_step * args.steps / (ms_total * 1e-3)
        e2e_value = pts_per_step * e2e_steps / (e2e_ms * 1e-3)
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
        pts_launch = SUB * POINTS_PER_FIELD              # one launch of a sweep kernel = one sub-batch
        gx = BYTES_SWEEP_X * pts_launch / (seg_ms[2] * 1e-3) / 1e9 if seg_ms[2] > 0 else 0.0
        gy = BYTES_SWEEP_Y * pts_launch / (seg_ms[3] * 1e-3) / 1e9 if seg_ms[3] > 0 else 0.0
        dominant_is_x = seg_ms[2] >= seg_ms[3]
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                tj = json.load(f)
            traffic = tj['sweep_x_bytes_per_point' if dominant_is_x else 'sweep_y_bytes_per_point'] * pts_launch
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': config_dict(F, world),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d * world, 'd2h_bytes_per_step': d2h * world,
                    'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps,
                    'api': 'fb_barnes_host (C ABI, pinned host buffers, H2D + kernels + D2H inside; every call returns with its fields in host memory)',
                    'host_cpus': numa,
                    # aggregate device -> pinned-host copy rate of the result buffers, all ranks copying at once (max time
                    # over ranks); the end-to-end arm cannot run faster than its D2H bytes at this rate
                    'host_ceiling_GBps': SUB * POINTS_PER_FIELD * 4 * world / (d2h_ms * 1e-3) / 1e9,
                    'host_ceiling_value': SUB * POINTS_PER_FIELD * world / (d2h_ms * 1e-3),
                    'frac_of_host_ceiling': e2e_value / (SUB * POINTS_PER_FIELD * world / (d2h_ms * 1e-3)),
                    'fields_per_call': EF, 'calls_per_step': ncalls},
            'gpu_launches': int(launches),
            'roofline': {
                'bound': 'hbm',
                'kernel': 'fb_sweepq_kernel<4,1,1> (x sweep: 4 fused passes per warp, TMA row staging, rings in tensor + shared memory, TMA transposing store)' if dominant_is_x
                          else 'fb_sweepq_kernel<4,1,2> (y sweep: 4 fused passes per warp + mask/divide/cast, TMA row staging, rings in tensor + shared memory)',
                'achieved': gx if dominant_is_x else gy, 'peak': peak, 'unit': 'GB/s',
                'frac': (gx if dominant_is_x else gy) / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_point': BYTES_SWEEP_X if dominant_is_x else BYTES_SWEEP_Y,
                'points_per_launch': pts_launch,
                'ms_per_launch': float(seg_ms[2] if dominant_is_x else seg_ms[3]),
                'other_kernels': {'sweep_x_GBps': gx, 'sweep_y_GBps': gy, 'sweep_x_frac': gx / peak, 'sweep_y_frac': gy / peak,
                                  'ms_zero_fill': float(seg_ms[0]),
                                  'ms_minmax_inject': float(seg_ms[1]), 'ms_sweep_x': float(seg_ms[2]),
                                  'ms_sweep_y': float(seg_ms[3])},
                'whole_step': {'algorithmic_bytes_per_point': BYTES_TOTAL_2D,
                               'achieved_GBps': BYTES_TOTAL_2D * F * POINTS_PER_FIELD * args.steps / (ms_total * 1e-3) / 1e9,
                               'frac_of_peak': BYTES_TOTAL_2D * F * POINTS_PER_FIELD * args.steps / (ms_total * 1e-3) / 1e9 / peak},
            },
            'clocks': clocks,
        }
        if fp32:
            s32 = fp32[1]
            bx, by = 16, 12     # fp32 sweeps: x reads and writes float2 nodes (8 + 8 B); y reads 8 B, writes the float32 field (4 B)
            line['fp32_path'] = {
                'note': 'same batches with FB_FLAG_FP32 (sweeps in fp32 working precision, tolerance-level parity); secondary number',
                'value': pts_per_step * args.steps / (ms32_total * 1e-3), 'unit': UNIT, 'ms_per_step': ms32_total / args.steps,
                'ms_zero_fill': float(s32[0]), 'ms_minmax_inject': float(s32[1]), 'ms_sweep_x': float(s32[2]),
                'ms_sweep_y': float(s32[3]),
                'sweep_x_GBps': bx * pts_launch / (s32[2] * 1e-3) / 1e9 if s32[2] > 0 else 0.0,
                'sweep_y_GBps': by * pts_launch / (s32[3] * 1e-3) / 1e9 if s32[3] > 0 else 0.0,
                'sweep_x_frac_of_peak': bx * pts_launch / (s32[2] * 1e-3) / 1e9 / peak if s32[2] > 0 else 0.0,
                'sweep_y_frac_of_peak': by * pts_launch / (s32[3] * 1e-3) / 1e9 / peak if s32[3] > 0 else 0.0}
        line.update(extra)
        if slab3d is not None:
            line['slab3d'] = slab3d
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            nf = max(cores, 8) * 6          # about 20 core-seconds of CPU work
            fps, secs, ref0 = cpu_fields_per_second(nf, cores)
            line['cpu_baseline'] = {'value': fps * POINTS_PER_FIELD, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': '%d fields of the same workload, one field per host thread, %.1f s wall'
                                              % (nf, secs)}
            try:
                with open(os.path.join(ROOT, 'profiles', 'r2_numba_reference_cpu.json')) as f:
                    nj = json.load(f)
                line['cpu_baseline']['numba_reference_single_thread'] = {
                    'value': nj['cases']['c5_field_N50000']['grid_points_per_s'], 'unit': UNIT, 'where': nj['where'],
                    'note': 'the unmodified reference under Numba, 1 thread, measured in the build container (its sources '
                            'are not in this repository); orientation only'}
            except Exception:
                pass
            # the oracle computed field 0 of sample set 0 (seed 2000) a moment ago: the timed arm and the end-to-end arm
            # must reproduce it bit for bit
            same_dev = bool(np.array_equal(check_dev.view(np.uint32), ref0.view(np.uint32)))
            same_e2e = bool(np.array_equal(check_e2e.view(np.uint32), ref0.view(np.uint32)))
            line['parity_checked'] = same_dev and same_e2e
            line['parity'] = {'field': 'seed 2000 (field 0 of sample set 0)', 'device_resident_arm_bit_identical': same_dev,
                              'e2e_arm_bit_identical': same_e2e, 'against': 'oracle port (oracle/fb_oracle.c), float32 field'}
        print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--fields', type=int, default=1024, help='fields per GPU per step (run as sub-batches of 64)')
    ap.add_argument('--no-extra', action='store_true', help='skip the single-field and S2 blocks')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--streams', type=int, default=2, help='device-resident arm: batches alternate over this many streams')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-fp32', action='store_true', help='skip the secondary fp32 working-precision measurement')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner with
    # printf), so file descriptor 1 is pointed at stderr and the JSON line goes to a duplicate of the
    # original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch one process per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               '--gpus', str(args.gpus), '--steps', str(args.steps), '--warmup', str(args.warmup),
               '--fields', str(args.fields), '--streams', str(args.streams)]
        cmd += [f for f, on in (('--no-extra', args.no_extra), ('--no-cpu', args.no_cpu), ('--no-fp32', args.no_fp32)) if on]
        sys.exit(subprocess.call(cmd, stdout=_JSON_OUT))     # the ranks' stdout is the original stdout
    run_gpu(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
