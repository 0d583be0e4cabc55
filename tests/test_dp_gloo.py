""" World-size-2 data-parallel host logic on CPU (gloo): flat gradient buckets, bucket partition in reverse parameter order, SUM all-reduce of every
bucket exactly once per step (hook-driven and `finish()` paths), per-rank seeds / DistributedSampler sharding. The CUDA kernels are not involved:
gradients are written into the flat buffer by hand, as the backward kernels would. """
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, hp):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deepcv_b200.meta.base_module import DeepcvModule
        from deepcv_b200.meta.flat_params import FlatAdamW, GradientBucketReducer, flatten_parameters
        from deepcv_b200.meta.ignite_training import BackendConfig, DataParallelModel
        torch.manual_seed(100 + rank)                       # different initial weights per rank: the wrap must broadcast rank 0's
        model = DeepcvModule((3, 32, 32), hp)
        dp = DataParallelModel(model, bucket_bytes=4 << 10)   # small buckets: several of them on the 68 KB default net
        flat = dp.flat
        assert len(flat.buckets) >= 3 and flat.buckets[0][1] == flat.numel and flat.buckets[-1][0] == 0
        covered = sorted(flat.buckets)
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))        # contiguous, non-overlapping
        assert all(p.grad is flat.grad_view(p) and p.grad.data_ptr() >= flat.flat_grads.data_ptr() for p in model.parameters())
        # parameters are views of the flat buffer and equal rank 0's after the broadcast
        gathered = [torch.empty_like(flat.flat_params) for _ in range(world)]
        dist.all_gather(gathered, flat.flat_params)
        assert torch.equal(gathered[0], gathered[1])
        # a fused layer's gradient targets point into the flat buffer, with the weight's physical [K][R][S][C] order
        layer = flat.layers[0]
        tgt = layer._grad_out['weight']
        assert tgt.shape == layer._op.weight.shape and tgt.permute(0, 2, 3, 1).is_contiguous()

        # ---- reduction: hooks path (fire every layer's hook in backward order) and finish() path
        for use_hooks in (True, False):
            dp.reducer.begin_step()
            flat.flat_grads.fill_(float(rank + 1))
            if use_hooks:
                for layer in reversed(flat.layers):
                    for hook in layer._backward_hooks.values():
                        hook(layer, None, None)
            dp.finish_gradient_reduction()
            assert torch.all(flat.flat_grads == 3.0), 'every bucket must be SUM-reduced exactly once'
        # the 1/world scaling is folded into the optimizer
        opt = FlatAdamW(model.parameters(), lr=1e-3).attach(flat)
        opt.grad_scale = 1. / dp.world_size
        assert opt.grad_scale == 0.5
        opt.zero_grad()                                     # must keep (not drop) the gradient views
        assert all(p.grad is flat.grad_view(p) for p in model.parameters())

        # sampler sharding + per-rank seed as in train()
        ds = torch.utils.data.TensorDataset(torch.arange(64))
        sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=False)
        idx = torch.tensor(list(iter(sampler)))
        both = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(both, idx)
        assert sorted(torch.cat(both).tolist()) == list(range(64))
        conf = BackendConfig('cpu', dist_backend='gloo', dist_url='env://')
        assert conf.distributed and conf.rank == rank and conf.gpus_world_size == world
    finally:
        dist.destroy_process_group()


def test_data_parallel_reducer_world_size_2(default_hp):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, default_hp), nprocs=2, join=True)
