""" World-size-2 data-parallel host logic on CPU (gloo): flat gradient buckets, bucket partition in reverse parameter order, SUM all-reduce of every
bucket exactly once per step (early launches driven by the layers' end-of-backward notifications, and the `finish()` path), gradient averaging,
per-rank seeds / DistributedSampler sharding. The CUDA kernels are not involved: gradients are written into the flat buffer by hand, as the backward
kernels would, and `ops._backward_done` is called as the end of each layer's backward does. """
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

        # ---- reduction: early launches (every layer reports the end of its backward, in backward order) and the finish() path
        from deepcv_b200 import ops
        assert all(not l._backward_hooks for l in flat.layers), 'module full-backward hooks fire before a first layer has enqueued its weight gradient'
        dp.reducer.average_in_finish = False
        for early in (True, False):
            dp.reducer.begin_step()
            flat.flat_grads.fill_(float(rank + 1))
            if early:
                launched_before_last = None
                for i, layer in enumerate(reversed(flat.layers)):
                    if i == len(flat.layers) - 1:
                        launched_before_last = list(dp.reducer._launched)
                    ops._backward_done(layer._grad_out, layer._grad_out)
                    assert layer._grad_out['_written']
                # the bucket holding the first layer's gradients is only reduced once that layer itself has reported
                first_buckets = dp.reducer._layer_buckets[id(flat.layers[0])]
                assert not any(launched_before_last[b] for b in first_buckets) and any(launched_before_last)
            dp.finish_gradient_reduction()
            assert torch.all(flat.flat_grads == 3.0), 'every bucket must be SUM-reduced exactly once'
        # averaging: by finish() (any optimizer) ...
        dp.reducer.average_in_finish = True
        flat.flat_grads.fill_(float(rank + 1))
        dp.finish_gradient_reduction()
        assert torch.all(flat.flat_grads == 1.5)
        # ... a second backward through one layer before its bucket has gone (shared weights): that bucket waits for finish()
        dp.reducer.average_in_finish = False
        flat.flat_grads.fill_(1.0)
        last = flat.layers[-1]
        ops._backward_done(last._grad_out, last._grad_out)
        dp.finish_gradient_reduction()
        assert torch.all(flat.flat_grads == 2.0)
        # ... or folded into the optimizer kernel
        opt = FlatAdamW(model.parameters(), lr=1e-3).attach(flat)
        opt.grad_scale = 1. / dp.world_size
        assert opt.grad_scale == 0.5
        flat.flat_grads.fill_(7.0)
        opt.zero_grad()                                     # keeps (does not drop) the gradient views, zeroes the buffer, re-arms direct writes
        assert all(p.grad is flat.grad_view(p) for p in model.parameters())
        assert torch.all(flat.flat_grads == 0) and not any(l._grad_out['_written'] for l in flat.layers)
        sd = opt.state_dict()
        assert sd['flat_step'] == 0

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
