""" Numerical test of the data-parallel path on real GPUs over NCCL (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single-GPU box). One process per GPU, DIFFERENT data on every rank, real backward kernels writing into the flat gradient buffer, the
bucketed all-reduce launched from the layers' end-of-backward notifications on the communication stream:

  * after backward + `finish_gradient_reduction()` the (averaged) gradients of every rank equal the MEAN of the per-rank oracle gradients
    (each rank computes its own with the CPU oracle; they are averaged with an all-reduce) — in particular those of the FIRST layer, whose
    weight-gradient kernels are enqueued after PyTorch would fire a module full-backward hook (round-1 advisor finding);
  * after several optimisation steps — eager and CUDA-graph replayed — the parameters of all ranks are BIT-identical.

Reference call site: `DistributedDataParallel(model, device_ids=[local_rank])`, /root/reference/src/deepcv/meta/ignite_training.py:373-390. """
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, hp, bucket_bytes):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    code = 0
    try:
        from deepcv_b200.meta.base_module import DeepcvModule
        from deepcv_b200.meta.flat_params import FlatAdamW
        from deepcv_b200.meta.ignite_training import CrossEntropyLoss, DataParallelModel, Engine, GraphedTrainStep, make_eager_process_function
        from oracle.deepcv_oracle import OracleDeepcvModule, train_step
        torch.manual_seed(7)                       # same oracle weights everywhere ...
        oracle = OracleDeepcvModule((3, 32, 32), hp)
        torch.manual_seed(100 + rank)              # ... different initial device weights per rank: the wrap must broadcast rank 0's
        model = DeepcvModule((3, 32, 32), hp).to(dev)
        if rank == 0:
            model.load_state_dict(oracle.state_dict())
        dp = DataParallelModel(model, bucket_bytes=bucket_bytes)
        flat = dp.flat
        if bucket_bytes < (1 << 20):
            assert len(flat.buckets) >= 3
        g = torch.Generator().manual_seed(1000 + rank)      # different data per rank
        batches = [(torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)) for _ in range(4)]
        loss_fn = CrossEntropyLoss()

        # ---- one backward: averaged device gradients == mean over ranks of the oracle's gradients
        x, y = batches[0]
        train_step(oracle, x, y)
        oracle64 = __import__('copy').deepcopy(oracle).double()
        train_step(oracle64, x.double(), y)
        dp.train()
        loss = loss_fn(dp(x.to(dev)), y.to(dev))
        flat.reset_gradients()
        loss.backward()
        dp.finish_gradient_reduction()              # average_in_finish: SUM all-reduce, then 1 / world
        torch.cuda.synchronize()
        names = [n for n, _ in model.named_parameters()]
        for (n, p), (_, p32), (_, p64) in zip(model.named_parameters(), oracle.named_parameters(), oracle64.named_parameters()):
            m32, m64 = p32.grad.to(dev), p64.grad.to(dev)
            dist.all_reduce(m32), dist.all_reduce(m64)
            m32, m64 = m32 / world, m64 / world
            allowed = 1e-4 * float(m64.abs().max()) + 8. * float((m32.double() - m64).abs().max()) + 1e-30
            err = float((p.grad.double() - m64).abs().max())
            assert err <= allowed, f'rank {rank}: gradient of {n} is not the mean of the per-rank gradients ({err / allowed:.2f}x the bound)'
        assert names[0].endswith('weight')          # the first layer's convolution weight is covered by the loop above
        del loss   # its autograd graph keeps the parameters' AccumulateGrad nodes (bound to this stream) alive: a later capture on another stream would depend on uncaptured work

        # ---- several optimisation steps, eager then graph-replayed: parameters bit-identical on all ranks
        def params_identical(tag):
            mine = flat.flat_params.clone()
            ref = mine.clone()
            dist.broadcast(ref, src=0)
            same = torch.tensor([int(torch.equal(mine, ref))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            assert int(same) == 1, f'{tag}: parameters differ between ranks'
        opt = FlatAdamW(model.parameters(), lr=1e-2, weight_decay=1e-2).attach(flat)
        opt.grad_scale = 1. / world
        dp.reducer.average_in_finish = False
        step = make_eager_process_function({}, dev, dp, {'main_loss': loss_fn}, opt)
        engine = Engine(step)
        before = flat.flat_params.clone()
        for x, y in batches[:2]:
            step(engine, (x, y))
        params_identical('eager')
        assert not torch.equal(before, flat.flat_params)
        runner = GraphedTrainStep(dp, loss_fn, opt, batches[0][0].to(dev), batches[0][1].to(dev), warmup_iters=2)
        params_identical('after capture')
        for x, y in batches[2:] + batches[:2]:
            runner.step(x.to(dev), y.to(dev))
        torch.cuda.synchronize()
        params_identical('graph replay')
        # BatchNorm statistics stay per replica (different data => different running means)
        rm = next(b for n, b in model.named_buffers() if n.endswith('running_mean')).clone()
        other = rm.clone()
        dist.broadcast(other, src=0)
        if rank == 1:
            assert not torch.equal(rm, other), 'BatchNorm running statistics must not be synchronised (per-replica BN)'
        if hasattr(runner, 'graph'):
            runner.graph.reset()
        torch.cuda.synchronize()
    except BaseException:
        import traceback
        traceback.print_exc()
        code = 1
    finally:
        # a process group whose communicator was captured into a CUDA graph can hang in destroy_process_group(): leave without the collective teardown
        sys.stdout.flush(), sys.stderr.flush()
        os._exit(code)


@pytest.mark.parametrize('bucket_bytes', [4 << 10, 8 << 20], ids=['several-buckets', 'one-bucket'])
def test_data_parallel_gradients_and_parameters_nccl(default_hp, bucket_bytes):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, default_hp, bucket_bytes), nprocs=2, join=True)


def _peer_worker(rank, world, port):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    code = 0
    try:
        from deepcv_b200.meta.flat_params import PeerAllReduce
        numel = 40000 + 3                                   # ragged: the last slice is not a whole number of 16-byte vectors
        buf = torch.zeros(numel, device=dev)
        peer = PeerAllReduce(buf, None)
        assert peer.world == world and peer.max_floats >= 17010
        g = torch.Generator().manual_seed(50 + rank)
        slices = [(0, 17016, 0), (17016, 20000, 1), (20000, numel, 2)]      # (start, end, slot): three "buckets"

        def reference(mine):
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            out = torch.zeros_like(mine)
            for p in parts:                                  # rank order 0..W-1: the kernel's summation order => bit-identical
                out += p
            return out
        for it in range(3):                                  # epochs advance; nothing is reset between calls
            mine = torch.randn(numel, generator=g).to(dev)
            buf.copy_(mine)
            ref = reference(mine)
            torch.cuda.synchronize(); dist.barrier()
            for start, end, slot in slices:
                peer.all_reduce(start, end, slot)
            torch.cuda.synchronize()
            assert torch.equal(buf, ref), f'rank {rank} iteration {it}: max |diff| {float((buf - ref).abs().max())}'
            dist.barrier()
        # CUDA-graph replay (the training step is one graph): static buffer, new data copied in before every replay
        static_in = torch.zeros(numel, device=dev)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            buf.copy_(static_in)
            for start, end, slot in slices:
                peer.all_reduce(start, end, slot)
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
        with torch.cuda.graph(graph):
            buf.copy_(static_in)
            for start, end, slot in slices:
                peer.all_reduce(start, end, slot)
        for it in range(3):
            mine = torch.randn(numel, generator=g).to(dev)
            static_in.copy_(mine)
            ref = reference(mine)
            torch.cuda.synchronize(); dist.barrier()
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(buf, ref), f'rank {rank} graph replay {it}: max |diff| {float((buf - ref).abs().max())}'
            dist.barrier()
        graph.reset()
        torch.cuda.synchronize()
    except BaseException:
        import traceback
        traceback.print_exc()
        code = 1
    finally:
        sys.stdout.flush(), sys.stderr.flush()
        os._exit(code)


def test_peer_memory_all_reduce_bit_exact():
    """ `dcv_peer_allreduce_sum` (one-shot all-reduce of small gradient buckets over NVLink peer memory) against the sum in rank order of all ranks'
    buffers: bit-identical on every rank, across repeated calls (epoch flags) and under CUDA-graph replay. """
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    mp.spawn(_peer_worker, args=(2, _free_port()), nprocs=2, join=True)
