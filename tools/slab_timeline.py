"""
Timeline of one z-slab call under torchrun (CUDA events on the main and the exchange stream, ms since the start of the call):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/slab_timeline.py [reserve_sms]
Shows whether the halo exchange overlaps with the sweeps of the interior planes.
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fast-barnes-py_b200'))
import numpy as np
import torch
import torch.distributed as dist
from fastbarnes import distributed as fd, _lib

rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
_lib.check(_lib.lib().fb_set_device(local))
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
reserve = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H, D, N, sigma = 1024, 1024, 512, 10_000_000, 8.0
rng = np.random.default_rng(1235)
pts = rng.uniform(0, 1, (N, 3)) * [W - 1, H - 1, D - 1]
val = rng.normal(0, 1, N)
dp = torch.from_numpy(pts).to(dev); dv = torch.from_numpy(val).to(dev)
mode = sys.argv[2] if len(sys.argv) > 2 else 'nccl'
slab = fd.BarnesSlab3D(sigma, [0.0] * 3, 1.0, (W, H, D), N, num_iter=4, reserve_sms=reserve, exchange=mode)
L = _lib.lib()
for _ in range(3):
    slab(dp, dv)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

main, comm = torch.cuda.current_stream(), slab.comm_stream
marks = []
if slab.exchange_mode == 'peer':
    # the peer-mapped exchange, step by step as BarnesSlab3D._call_peer does it
    import time
    t0 = ev(); h0 = time.perf_counter()
    slab._host_barrier()
    for q, send, recv in slab.transfers():
        if send:
            slab.peers[q]['pulled'].wait(main)
    slab.inject(dp, dv); marks.append(('inject done', ev()))
    lo, hi = slab.lo_need, slab.hi_need
    slab.sweeps(0, lo); slab.ev_down.record(main); marks.append(('low boundary swept', ev()))
    slab.sweeps(slab.zc - hi, hi); slab.ev_up.record(main); marks.append(('high boundary swept', ev()))
    h1 = time.perf_counter()
    slab._host_barrier()
    h2 = time.perf_counter()
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        marks.append(('pulls start', ev())); slab._pull('down'); marks.append(('down pulled', ev()))
        slab._pull('up'); marks.append(('up pulled', ev())); slab.ev_pulled.record(comm)
    slab.sweeps(lo, slab.zc - lo - hi); marks.append(('interior swept', ev()))
    main.wait_stream(comm)
    slab.phase2(); marks.append(('z sweep done', ev()))
    torch.cuda.synchronize()
    line = {'rank': rank, 'mode': 'peer', 'host_ms_enqueue_until_barrier': round((h1 - h0) * 1e3, 3), 'host_ms_barrier': round((h2 - h1) * 1e3, 3),
            'ms': {k: round(t0.elapsed_time(e), 3) for k, e in marks}}
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        for o in out:
            print(json.dumps(o))
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
t0 = ev()
slab.inject(dp, dv); marks.append(('inject done', ev()))
lo, hi = slab.lo_need, slab.hi_need
slab.sweeps(0, lo); marks.append(('low boundary swept', ev()))
comm.wait_stream(main)
with torch.cuda.stream(comm):
    marks.append(('down exchange starts', ev())); slab.exchange('down'); marks.append(('down exchange done', ev()))
L.fb_set_option(b'sweepq_reserve_sms', reserve)
slab.sweeps(slab.zc - hi, hi); marks.append(('high boundary swept', ev()))
comm.wait_stream(main)
with torch.cuda.stream(comm):
    marks.append(('up exchange starts', ev())); slab.exchange('up'); marks.append(('up exchange done', ev()))
slab.sweeps(lo, slab.zc - lo - hi); marks.append(('interior swept', ev()))
L.fb_set_option(b'sweepq_reserve_sms', 0)
main.wait_stream(comm)
slab.phase2(); marks.append(('z sweep done', ev()))
torch.cuda.synchronize()
line = {'rank': rank, 'reserve_sms': reserve, 'ms': {k: round(t0.elapsed_time(e), 3) for k, e in marks}}
out = [None] * world
dist.all_gather_object(out, line)
if rank == 0:
    for o in out:
        print(json.dumps(o))
dist.barrier(); dist.destroy_process_group()
