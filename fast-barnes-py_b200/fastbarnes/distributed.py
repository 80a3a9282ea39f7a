# -*- coding: utf-8 -*-
"""
Multi-GPU use of the batched path: independent fields (time steps, ensemble members) are
partitioned over the ranks of a torch.distributed process group, one process per GPU.  The path has
no exchange step, so there is NO data-path collective: every rank interpolates its own contiguous
block of fields.  An optional all-gather assembles the full result on every rank (host tensors
over gloo, device tensors over NCCL).

The reference has no counterpart (it is single-threaded); this module only adds plumbing around
`interpolation.barnes_batched`.
"""
import numpy as np

from . import interpolation


def shard_range(nfields, world_size, rank):
    """ Contiguous, balanced block [b0, b1) of `nfields` fields owned by `rank` (first ranks get the remainder). """
    if world_size < 1 or not (0 <= rank < world_size):
        raise RuntimeError('invalid rank / world size: %d / %d' % (rank, world_size))
    base, rem = divmod(int(nfields), int(world_size))
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_samples(sample_offsets, b0, b1):
    """ Sample rows [s0, s1) of the fields [b0, b1) and their offsets rebased to 0. """
    sample_offsets = np.asarray(sample_offsets, dtype=np.int64)
    s0, s1 = int(sample_offsets[b0]), int(sample_offsets[b1])
    return s0, s1, sample_offsets[b0:b1 + 1] - s0


def barnes_batched_sharded(pts, val, sigma, x0, step, size, sample_offsets, method='optimized_convolution',
                           num_iter=4, max_dist=3.5, group=None, gather=True, compute=None):
    """
    Every rank of `group` (default: the world group; without an initialised process group this is a
    plain single-process call) interpolates its block of the B fields described by
    (`pts`, `val`, `sample_offsets`) -- see `interpolation.barnes_batched`.

    gather=True: returns the full float32 array (B,) + size[::-1] on every rank.
    gather=False: returns (b0, local) with `local` the rank's fields [b0, b0 + len(local)).
    `compute` is the per-rank batched interpolation, by default `interpolation.barnes_batched`
    (the CUDA path); tests substitute a checker.
    """
    import torch
    import torch.distributed as dist
    compute = interpolation.barnes_batched if compute is None else compute
    sample_offsets = np.asarray(sample_offsets, dtype=np.int64)
    nfields = len(sample_offsets) - 1
    active = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if active else 1
    rank = dist.get_rank(group) if active else 0
    b0, b1 = shard_range(nfields, world, rank)
    rsize = tuple(int(s) for s in (size if isinstance(size, (list, tuple, np.ndarray)) else (size,)))[::-1]
    if b1 > b0:
        s0, s1, offs = shard_samples(sample_offsets, b0, b1)
        local = compute(pts[s0:s1], val[s0:s1], sigma, x0, step, size, sample_offsets=offs, method=method,
                        num_iter=num_iter, max_dist=max_dist)
        local = np.ascontiguousarray(local, dtype=np.float32).reshape((b1 - b0,) + rsize)
    else:
        local = np.empty((0,) + rsize, dtype=np.float32)
    if not gather:
        return b0, local
    if world == 1:
        return local
    # all ranks need equally shaped tensors: pad every block to the largest one
    per = [shard_range(nfields, world, r) for r in range(world)]
    nmax = max(e - b for b, e in per)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.zeros((nmax,) + rsize, dtype=torch.float32, device=dev)
    if b1 > b0:
        mine[:b1 - b0] = torch.from_numpy(local).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty((nfields,) + rsize, dtype=np.float32)
    for r, (b, e) in enumerate(per):
        if e > b:
            out[b:e] = parts[r][:e - b].cpu().numpy()
    return out


# ---------------------------------------------------------------------------------------------
# 3D volumes: z-slab decomposition with halo exchange

def slab_transfers(nplanes, world, rank, halo):
    """ Halo exchange plan of rank `rank` for `nplanes` z planes split over `world` ranks (shard_range) with a halo of
    `halo` planes on each side: [(peer, send (a, b) or None, recv (a, b) or None)] in absolute plane numbers, `send` =
    own planes inside the peer's extended window, `recv` = planes of the peer inside the own extended window.  The
    windows are cut at the ends of the volume and may span several ranks (slabs thinner than the halo). """
    ranges = [shard_range(nplanes, world, r) for r in range(world)]
    z0, z1 = ranges[rank]
    e0, e1 = max(0, z0 - halo), min(nplanes, z1 + halo)
    out = []
    for q, (q0, q1) in enumerate(ranges):
        if q == rank:
            continue
        qe0, qe1 = max(0, q0 - halo), min(nplanes, q1 + halo)
        s0, s1 = max(z0, qe0), min(z1, qe1)
        r0, r1 = max(q0, e0), min(q1, e1)
        send = (s0, s1) if s1 > s0 else None
        recv = (r0, r1) if r1 > r0 else None
        if send or recv:
            out.append((q, send, recv))
    return out


class _IpcEvent:
    """ CUDA event shared between the processes of one node (C ABI fb_ipc_event_*; torch's Event.from_ipc_handle objects
    crashed in wait() with torch 2.11).  Owner: create(); others: open(handle). """

    def __init__(self, lib_module, handle=None):
        import ctypes
        self._lib = lib_module
        self.ptr = ctypes.c_void_p()
        L = lib_module.lib()
        if handle is None:
            buf = ctypes.create_string_buffer(64)
            lib_module.check(L.fb_ipc_event_create(ctypes.byref(self.ptr), buf))
            self.handle = buf.raw
        else:
            self.handle = bytes(handle)
            lib_module.check(L.fb_ipc_event_open(self.handle, ctypes.byref(self.ptr)))

    def record(self, stream):
        self._lib.check(self._lib.lib().fb_event_record(self.ptr, stream.cuda_stream))

    def wait(self, stream):
        """ work submitted to `stream` from now on waits for the most recent record (as of this call) """
        self._lib.check(self._lib.lib().fb_stream_wait_event(stream.cuda_stream, self.ptr))


class BarnesSlab3D:
    """
    3D optimized-convolution Barnes interpolation of ONE large volume split into z-slabs, one per
    rank (one process per GPU).  Every rank sees all samples (they are small next to the volume),
    injects and x/y-sweeps only its own planes -- bit-identical to the single-GPU planes, since
    those sweeps are independent per plane -- then receives the `halo = num_iter*(T_z+1)` planes
    below and above its slab from the ranks that own them (torch.distributed send/recv: NCCL over
    NVLink on GPUs; the planes may come from more than one rank when slabs are thinner than the halo)
    and runs the fused z sweep + mask + divide + cast over its extended lines.

    The exchange is overlapped with the sweeps: the planes other ranks need are swept first, their
    transfer runs on a second stream while the interior planes are swept.

    Parity against the single-GPU run: within the halo every input an own plane depends on is present;
    the result differs only by where the sliding accumulator of the z sweep starts, i.e. by rounding:
    |difference of the fp64 quotient| <= 1e-12 * (range of the values), identical NaN mask.  (A bound
    RELATIVE to the quotient itself is meaningless where the field crosses zero.)

    `nslabs`/`slab` may be given explicitly to run several slabs one after the other in a single
    process (used by the tests to check the decomposition on one GPU).
    """

    def __init__(self, sigma, x0, step, size, nsamples, method='optimized_convolution', num_iter=4, max_dist=3.5,
                 group=None, device=None, want_float64=False, nslabs=None, slab=None, reserve_sms=0, exchange='nccl'):
        import torch
        import torch.distributed as dist
        from . import _lib
        from .interpolation import _per_axis, _grid_size, _problem, _check_kernel_vs_grid, _CONV_METHODS
        from math import exp
        if method not in _CONV_METHODS:
            raise RuntimeError("encountered invalid Barnes interpolation method: " + str(method))
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device available; this package has no CPU fallback')
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.group = group
        active = dist.is_available() and dist.is_initialized()
        self.world = int(nslabs) if nslabs is not None else (dist.get_world_size(group) if active else 1)
        self.rank = int(slab) if slab is not None else (dist.get_rank(group) if active else 0)
        self.use_dist = active and nslabs is None and self.world > 1
        sigma = _per_axis('sigma', sigma, 3)
        x0 = _per_axis('x0', x0, 3)
        step = _per_axis('step', step, 3)
        self.size = _grid_size(size, 3)
        _check_kernel_vs_grid(method, sigma, step, self.size, num_iter)
        self.prob = _problem(3, sigma, x0, step, self.size, _CONV_METHODS[method], num_iter, exp(-max_dist ** 2 / 2), 1)
        L = _lib.lib()
        self.halo = int(L.fb_slab_halo_planes(self.prob))
        self.nsamples = int(nsamples)
        Dz = self.size[2]
        if self.world > Dz:
            raise RuntimeError('more ranks (%d) than z planes (%d)' % (self.world, Dz))
        self.ranges = [shard_range(Dz, self.world, r) for r in range(self.world)]
        self.z0, self.z1 = self.ranges[self.rank]
        # extended window: the halo below / above, cut at the ends of the volume (it may span several ranks)
        self.ext0, self.ext1 = max(0, self.z0 - self.halo), min(Dz, self.z1 + self.halo)
        self.halo_lo, self.halo_hi = self.z0 - self.ext0, self.ext1 - self.z1
        self.zc = self.z1 - self.z0
        self.z_ext = self.zc + self.halo_lo + self.halo_hi
        nb = _lib.ctypes.c_int64()
        ov = _lib.ctypes.c_int64()
        ow = _lib.ctypes.c_int64()
        _lib.check(L.fb_slab_layout(self.prob, self.nsamples, self.zc, self.halo_lo, self.halo_hi, int(want_float64),
                                    _lib.ctypes.byref(nb), _lib.ctypes.byref(ov), _lib.ctypes.byref(ow)))
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        W, H = self.size[0], self.size[1]
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
            nbytes = self.z_ext * H * W * 8
            il = int(L.fb_slab_interleaved(self.prob))
            if il < 0:
                _lib.check(il)
            self.interleaved = il == 1
            if self.interleaved:
                # one array of (value, weight) nodes; what the exchange moves are planes of it
                self.nodes = self.workspace[ov.value:ov.value + 2 * nbytes].view(torch.float64).view(self.z_ext, H, W, 2)
                self.vB, self.wB = self.nodes[..., 0], self.nodes[..., 1]
                self.planes = (self.nodes,)
            else:
                self.vB = self.workspace[ov.value:ov.value + nbytes].view(torch.float64).view(self.z_ext, H, W)
                self.wB = self.workspace[ow.value:ow.value + nbytes].view(torch.float64).view(self.z_ext, H, W)
                self.planes = (self.vB, self.wB)
            # the result stays where the z sweep writes it (the extended volume inside the workspace); `out` / `out64` are
            # views of the own planes, valid until the next call
            o32 = _lib.ctypes.c_int64()
            o64 = _lib.ctypes.c_int64()
            _lib.check(L.fb_slab_result_offsets(self.prob, self.nsamples, self.zc, self.halo_lo, self.halo_hi, int(want_float64),
                                                _lib.ctypes.byref(o32), _lib.ctypes.byref(o64)))
            a, b = self.halo_lo, self.halo_lo + self.zc
            self.out = self.workspace[o32.value:o32.value + nbytes // 2].view(torch.float32).view(self.z_ext, H, W)[a:b]
            self.out64 = (self.workspace[o64.value:o64.value + nbytes].view(torch.float64).view(self.z_ext, H, W)[a:b]
                          if want_float64 else None)
            self.comm_stream = torch.cuda.Stream(device=self.device) if self.use_dist else None
        self.want64 = bool(want_float64)
        # SMs the sweeps leave free while a halo exchange is in flight.  The sweep kernels are persistent, one CTA per SM
        # holding all of its shared memory, so the NCCL kernels of the exchange only start where SMs are left free -- and
        # they need many: measured on 2 B200 (tools/slab_timeline.py, profiles/r2_slab_timeline.json) the exchange runs
        # beside the interior sweeps with 48 SMs reserved, not with 16, and the sweeps then lose what the overlap gains.
        # Default 0: the exchange is enqueued early and effectively runs when the sweeps have drained.
        self.reserve_sms = int(reserve_sms)
        # planes of mine that other ranks need: [z0, z0 + lo_need) and [z1 - hi_need, z1)
        self.lo_need = min(self.zc, self.halo) if self.rank > 0 else 0
        self.hi_need = min(self.zc, self.halo) if self.rank < self.world - 1 else 0
        # transport of the halo planes: 'nccl' (send / recv kernels) or 'peer' (the ranks of one node map each other's
        # buffers -- CUDA IPC -- and PULL the planes with device-to-device copies on the copy engines, which need no SM and
        # therefore run beside the persistent sweep kernels; ordering by interprocess events)
        self.exchange_mode = 'nccl'
        if self.use_dist and exchange == 'peer':
            self._setup_peer_exchange()

    # -- peer-mapped exchange ------------------------------------------------------------------------
    def _setup_peer_exchange(self):
        """ Collective over the group: every rank publishes an IPC handle of its plane buffers and two interprocess
        events, opens those of the ranks it pulls from, and all agree on whether the mapping worked everywhere. """
        torch, dist = self.torch, self.dist
        from torch.multiprocessing.reductions import reduce_tensor
        ok = 1
        try:
            with torch.cuda.device(self.device):
                self.ev_down = _IpcEvent(self._lib)      # my lowest planes are swept
                self.ev_up = _IpcEvent(self._lib)        # my highest planes are swept
                self.ev_pulled = _IpcEvent(self._lib)    # I have read my peers' planes
                for e in (self.ev_down, self.ev_up, self.ev_pulled):
                    e.record(torch.cuda.current_stream())
                mine = {'rank': self.rank, 'device': self.device.index, 'ext0': self.ext0,
                        'planes': [reduce_tensor(b) for b in self.planes],
                        'events': [e.handle for e in (self.ev_down, self.ev_up, self.ev_pulled)]}
        except Exception as e:                           # no IPC in this environment
            mine, ok = {'rank': self.rank, 'error': repr(e)}, 0
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self.peers = {}
        if ok and all('error' not in m for m in everyone):
            try:
                need = {q for q, send, recv in self.transfers()}
                for q in sorted(need):
                    m = everyone[q]
                    planes = [fn(*args) for fn, args in m['planes']]
                    with torch.cuda.device(self.device):
                        evs = [_IpcEvent(self._lib, h) for h in m['events']]
                    self.peers[q] = {'ext0': m['ext0'], 'planes': planes, 'down': evs[0], 'up': evs[1], 'pulled': evs[2]}
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            # host-side rendezvous (the order of event records and waits across processes is an order of host calls)
            self.host_group = dist.new_group(backend='gloo')
            self.exchange_mode = 'peer'
            self.calls = 0
        else:
            self.peers = {}

    def _host_barrier(self):
        self.dist.barrier(group=self.host_group)

    def _pull(self, direction):
        """ copies, on the current stream, the planes that travel in `direction` from the ranks that own them into my
        halo: 'down' = the lowest planes of the ranks above me, 'up' = the highest planes of the ranks below me """
        torch = self.torch
        cur = torch.cuda.current_stream()
        for q, send, recv in self.transfers():
            if not recv:
                continue
            d = 'up' if q < self.rank else 'down'
            if d != direction:
                continue
            peer = self.peers[q]
            peer[d].wait(cur)                            # the owner has swept those planes (its record precedes the host barrier)
            for dst, src in zip(self.planes, peer['planes']):
                dst[recv[0] - self.ext0:recv[1] - self.ext0].copy_(src[recv[0] - peer['ext0']:recv[1] - peer['ext0']],
                                                                   non_blocking=True)

    def _call_peer(self, pts, val):
        torch = self.torch
        with torch.cuda.device(self.device):
            main, comm = torch.cuda.current_stream(), self.comm_stream
            # the ranks that pull from me must have finished reading the planes of the previous call before I overwrite them
            self._host_barrier()
            for q, send, recv in self.transfers():
                if send:
                    self.peers[q]['pulled'].wait(main)
            self.inject(pts, val)
            lo, hi = self.lo_need, self.hi_need
            if lo + hi >= self.zc:
                self.sweeps(0, lo)
                self.ev_down.record(main)
                self.sweeps(lo, self.zc - lo)
                self.ev_up.record(main)
            else:
                self.sweeps(0, lo)
                self.ev_down.record(main)
                self.sweeps(self.zc - hi, hi)
                self.ev_up.record(main)
            self._host_barrier()                         # every rank has issued its records of this call
            comm.wait_stream(main)                       # (also behind the z sweep of the previous call, which read the halo)
            with torch.cuda.stream(comm):
                self._pull('down')
                self._pull('up')
                self.ev_pulled.record(comm)
            if lo + hi < self.zc:
                self.sweeps(lo, self.zc - lo - hi)       # the interior, while the copy engines move the halos
            main.wait_stream(comm)
        return self.phase2()

    # -- who sends what to whom --------------------------------------------------------------------
    def transfers(self):
        """ the halo exchange plan of this rank (slab_transfers) """
        return slab_transfers(self.size[2], self.world, self.rank, self.halo)

    # -- the steps ---------------------------------------------------------------------------------
    def _check_inputs(self, pts, val):
        torch = self.torch
        if pts.dtype != torch.float64 or val.dtype != torch.float64 or not pts.is_cuda or not val.is_cuda:
            raise RuntimeError('pts and val must be float64 CUDA tensors')

    def phase1(self, pts, val):
        """ centring + injection + x / y sweeps of all own planes """
        L, _lib = self._lib.lib(), self._lib
        self._check_inputs(pts, val)
        with self.torch.cuda.device(self.device):
            st = self.torch.cuda.current_stream().cuda_stream
            _lib.check(L.fb_slab_phase1_dev(self.prob, self.z0, self.zc, self.halo_lo, self.halo_hi, self.nsamples,
                                            pts.data_ptr(), val.data_ptr(), int(self.want64),
                                            self.workspace.data_ptr(), self.workspace.numel(), st))

    def inject(self, pts, val):
        L, _lib = self._lib.lib(), self._lib
        self._check_inputs(pts, val)
        with self.torch.cuda.device(self.device):
            st = self.torch.cuda.current_stream().cuda_stream
            _lib.check(L.fb_slab_inject_dev(self.prob, self.z0, self.zc, self.halo_lo, self.halo_hi, self.nsamples,
                                            pts.data_ptr(), val.data_ptr(), int(self.want64),
                                            self.workspace.data_ptr(), self.workspace.numel(), st))

    def sweeps(self, plane_begin, plane_count):
        """ x / y sweeps of the own planes [plane_begin, plane_begin + plane_count) (relative to z0) """
        if plane_count <= 0:
            return
        L, _lib = self._lib.lib(), self._lib
        with self.torch.cuda.device(self.device):
            st = self.torch.cuda.current_stream().cuda_stream
            _lib.check(L.fb_slab_sweeps_dev(self.prob, self.z0, self.zc, self.halo_lo, self.halo_hi, self.nsamples,
                                            int(self.want64), int(plane_begin), int(plane_count),
                                            self.workspace.data_ptr(), self.workspace.numel(), st))

    def own_planes(self):
        """ (values, weights) views of the own planes inside the extended slab. """
        a, b = self.halo_lo, self.halo_lo + self.zc
        return self.vB[a:b], self.wB[a:b]

    def exchange(self, direction=None):
        """ halo exchange with the ranks whose planes lie in the extended window (torch.distributed point-to-point,
        one batch: NCCL groups the sends and receives).  Runs on the current stream.
        direction: None = everything; 'down' = the planes that travel to lower ranks (my lowest planes out, the lowest
        planes of the ranks above me in); 'up' = the planes that travel to higher ranks.  Every rank must call the same
        sequence of directions. """
        if not self.use_dist:
            return
        dist = self.dist
        ops = []
        for q, send, recv in self.transfers():
            send_dir = 'down' if q < self.rank else 'up'            # my planes travel towards q
            recv_dir = 'up' if q < self.rank else 'down'            # q's planes travel towards me
            for buf in self.planes:
                if send and direction in (None, send_dir):
                    ops.append(dist.P2POp(dist.isend, buf[send[0] - self.ext0:send[1] - self.ext0], q, self.group))
                if recv and direction in (None, recv_dir):
                    ops.append(dist.P2POp(dist.irecv, buf[recv[0] - self.ext0:recv[1] - self.ext0], q, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def phase2(self):
        torch, L, _lib = self.torch, self._lib.lib(), self._lib
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.fb_slab_phase2_inplace_dev(self.prob, self.z0, self.zc, self.halo_lo, self.halo_hi, self.nsamples,
                                                    int(self.want64), self.workspace.data_ptr(), self.workspace.numel(), st))
        return self.out

    def __call__(self, pts, val):
        """ pts (N, 3), val (N,) float64 CUDA tensors holding ALL samples; returns the rank's planes
        [z0, z1) as a float32 CUDA tensor (z1 - z0, H, W). """
        if not self.use_dist:
            self.phase1(pts, val)
            return self.phase2()
        if self.exchange_mode == 'peer':
            return self._call_peer(pts, val)
        torch = self.torch
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            comm = self.comm_stream
            L = self._lib.lib()
            self.inject(pts, val)
            # the planes the lower ranks need first; they travel ('down') on the second stream while the rest is swept, then
            # the planes the higher ranks need ('up') while the interior is swept
            lo, hi = self.lo_need, self.hi_need
            if lo + hi >= self.zc:
                # thin slab: the two boundary regions overlap -- lower part, send down, upper part, send up
                lo = min(lo, self.zc)
                self.sweeps(0, lo)
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    self.exchange('down')
                L.fb_set_option(b'sweepq_reserve_sms', self.reserve_sms)
                try:
                    self.sweeps(lo, self.zc - lo)
                finally:
                    L.fb_set_option(b'sweepq_reserve_sms', 0)
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    self.exchange('up')
            else:
                self.sweeps(0, lo)
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    self.exchange('down')
                L.fb_set_option(b'sweepq_reserve_sms', self.reserve_sms)
                try:
                    self.sweeps(self.zc - hi, hi)
                    comm.wait_stream(main)
                    with torch.cuda.stream(comm):
                        self.exchange('up')
                    self.sweeps(lo, self.zc - lo - hi)
                finally:
                    L.fb_set_option(b'sweepq_reserve_sms', 0)
            main.wait_stream(comm)
        return self.phase2()


def barnes_slabs_emulated(pts, val, sigma, x0, step, size, nslabs, num_iter=4, max_dist=3.5,
                          method='optimized_convolution', want_float64=False):
    """
    Runs the z-slab decomposition with `nslabs` slabs one after the other on the current GPU,
    copying the halo planes between the slabs' buffers (what the NCCL exchange does between ranks),
    with the same split of phase 1 (boundary planes first) the multi-GPU run uses.
    Returns the assembled float32 volume (and the fp64 quotient if requested) as numpy arrays.
    """
    import torch
    dp = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float64)).cuda()
    dv = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float64)).cuda()
    slabs = [BarnesSlab3D(sigma, x0, step, size, len(val), method=method, num_iter=num_iter, max_dist=max_dist,
                          want_float64=want_float64, nslabs=nslabs, slab=r) for r in range(nslabs)]
    for s in slabs:
        s.inject(dp, dv)
        lo, hi = s.lo_need, s.hi_need
        if lo + hi >= s.zc:
            s.sweeps(0, s.zc)
        else:
            s.sweeps(0, lo)
            s.sweeps(s.zc - hi, hi)
            s.sweeps(lo, s.zc - lo - hi)
    for s in slabs:
        for q, send, recv in s.transfers():
            if recv:
                src = slabs[q]
                for dst_buf, src_buf in zip(s.planes, src.planes):
                    dst_buf[recv[0] - s.ext0:recv[1] - s.ext0].copy_(src_buf[recv[0] - src.ext0:recv[1] - src.ext0])
    outs = [s.phase2() for s in slabs]
    torch.cuda.synchronize()
    vol = torch.cat(outs, dim=0).cpu().numpy()
    if want_float64:
        return vol, torch.cat([s.out64 for s in slabs], dim=0).cpu().numpy()
    return vol
