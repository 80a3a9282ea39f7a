""" Probe of the pieces the peer-mapped halo exchange needs, one at a time, under torchrun (2+ ranks of one node):
CUDA IPC mapping of a peer's tensor, a device-to-device pull from it, interprocess events.  Prints what works. """
import faulthandler, os, sys
faulthandler.enable()
import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
def say(*a):
    print('[rank %d]' % rank, *a, flush=True)

big = torch.empty(64 << 20, dtype=torch.uint8, device=dev)            # like the workspace: a view of a larger allocation
mine = big[1 << 20:(1 << 20) + 8 * 1024 * 1024].view(torch.float64).view(16, 256, 256)
mine.fill_(float(rank + 1))
torch.cuda.synchronize()
say('step 1: reduce_tensor')
red = reduce_tensor(mine)
say('step 2: all_gather_object')
allred = [None] * world
dist.all_gather_object(allred, {'device': local, 'red': red})
q = (rank + 1) % world
say('step 3: rebuild tensor of rank', q)
fn, args = allred[q]['red']
peer = fn(*args)
say('   peer tensor on', peer.device, tuple(peer.shape), 'first value', float(peer.reshape(-1)[0].item()))
say('step 4: pull copy')
dst = torch.zeros_like(mine)
dst[2:6].copy_(peer[2:6], non_blocking=True)
torch.cuda.synchronize()
say('   pulled value', float(dst[3, 0, 0].item()), 'expected', float(q + 1))
say('step 5: interprocess events')
ev = torch.cuda.Event(enable_timing=False, interprocess=True)
ev.record()
h = ev.ipc_handle()
hs = [None] * world
dist.all_gather_object(hs, {'device': local, 'h': h})
pev = torch.cuda.Event.from_ipc_handle(torch.device('cuda', hs[q]['device']), hs[q]['h'])
torch.cuda.current_stream().wait_event(pev)
torch.cuda.synchronize()
say('   waited on the event of rank', q)
say('step 6: gloo group')
g = dist.new_group(backend='gloo')
dist.barrier(group=g)
say('all steps ok')
dist.barrier(); dist.destroy_process_group()
