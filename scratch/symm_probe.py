""" Probe: torch symmetric memory on this box (peer pointers for a hand-written one-shot all-reduce). torchrun --nproc-per-node 2 scratch/symm_probe.py """
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
t.fill_(rank + 1)
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'ptrs', [hex(p) for p in h.buffer_ptrs], 'signal', [hex(p) for p in h.signal_pad_ptrs], h.signal_pad_size, 'multicast', h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None, flush=True)
dist.barrier(); torch.cuda.synchronize()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, 'peer value', float(peer[0]), flush=True)
# can a sub-range view / second allocation be made? capture-safety: pointers are stable
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g):
        t.add_(1.)
g.replay(); torch.cuda.synchronize(); dist.barrier()
print(rank, 'after graph', float(t[0]), float(peer[0]), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
