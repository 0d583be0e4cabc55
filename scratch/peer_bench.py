""" Latency of the one-shot peer all-reduce vs NCCL for small buffers (torchrun --nproc-per-node N scratch/peer_bench.py). """
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
from deepcv_b200.meta.flat_params import PeerAllReduce
for numel in (328, 4096, 17016, 65536, 131072):
    buf = torch.zeros(numel, device=dev)
    peer = PeerAllReduce(buf, None)
    plain = torch.zeros(numel, device=dev)
    res = {}
    for name, fn in (('peer', lambda: peer.all_reduce(0, numel, 0)), ('nccl', lambda: dist.all_reduce(plain))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g):
                for _ in range(50):
                    fn()
        torch.cuda.synchronize(); dist.barrier()
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) * 1e3 / 500
        dist.barrier()
    if rank == 0:
        print(f'{numel * 4 / 1024:8.1f} KB  peer {res["peer"]:6.2f} us  nccl {res["nccl"]:6.2f} us', flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
