""" Flat parameter / gradient buffers, the AdamW step over them, and the bucketed data-parallel gradient reducer.

Reference call sites this stands in for: `optimizer = opt(model.parameters(), **hp['optimizer_opts'])` with `opt = torch.optim.AdamW`
(`classification/image.py:71`, `meta/ignite_training.py:223`), `optimizer.zero_grad()` / `optimizer.step()` in the training step
(`ignite_training.py:252-254`) and `DistributedDataParallel(model, device_ids=[local_rank])` (`ignite_training.py:380`).

Layout in HBM: ONE fp32 buffer for all parameters and ONE for all gradients, in `model.parameters()` order, every tensor 16-byte
aligned. Each `torch.nn.Parameter` becomes a view of its slice (convolution weights keep their physical [K][R][S][C] order), each
`.grad` a view of the matching gradient slice, and the fused layers' backward kernels write parameter gradients *directly* into those
slices (`FusedLayer._grad_out`) — there is no per-parameter gradient tensor, no `zero_grad` pass and no flatten / unflatten copy. The
gradient buffer is cut into buckets (default 8 MB) in reverse parameter order = the order backward produces them; a bucket is
all-reduced (NCCL, side stream) as soon as its last producer has been enqueued, overlapping the rest of backward.
"""
import ctypes
import os
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from .._lib import check, lib
from .nn import FusedLayer

__all__ = ['FlatParameters', 'flatten_parameters', 'FlatAdamW', 'GradientBucketReducer']


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParameters:
    """ Flat fp32 `params` / `grads` buffers of a model whose parameters (and `.grad`s) have been re-pointed into them. """

    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 8 << 20):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError('flatten_parameters: model has no trainable parameter')
        if any(p.dtype != torch.float32 for p in params):
            raise TypeError('flatten_parameters: master parameters must be float32')
        device = params[0].device
        self.model, self.params = model, params
        self.offsets: List[int] = []
        total = 0
        for p in params:
            self.offsets.append(total)
            total += (p.numel() + 7) // 8 * 8  # 32-byte aligned fp32 slices = 16-byte aligned slices of a bf16 shadow of the buffer (TMA needs 16)
        self.numel = total
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=device)
        self.zeroed_by_step = False   # True while a captured step zeroes `flat_grads` itself (GraphedTrainStep)
        self._grad_views: Dict[int, torch.Tensor] = {}
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                view_p, view_g = self._views_like(p, off)
                view_p.copy_(p.data)
                p.data = view_p
                p.grad = view_g
                self._grad_views[id(p)] = view_g
        # buckets over the flat gradient buffer, in reverse parameter order (the order backward fills them)
        self.buckets: List[Tuple[int, int]] = []   # (start, end) element ranges
        self.bucket_of: Dict[int, int] = {}        # id(param) -> bucket index
        end = total
        limit = max(1, bucket_bytes // 4)
        for p, off in reversed(list(zip(params, self.offsets))):
            self.bucket_of[id(p)] = len(self.buckets)
            if end - off >= limit:
                self.buckets.append((off, end))
                end = off
        if end > 0:
            self.buckets.append((0, end))
        self._wire_layers()

    def _views_like(self, p: torch.Tensor, off: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """ Views of the two flat buffers with `p`'s shape and (dense) physical element order. """
        n = p.numel()
        if p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous() and not p.is_contiguous():
            k, c, r, s = p.shape
            mk = lambda buf: buf[off:off + n].view(k, r, s, c).permute(0, 3, 1, 2)
        else:
            if not p.is_contiguous():
                p.data = p.data.contiguous()
            mk = lambda buf: buf[off:off + n].view(p.shape)
        return mk(self.flat_params), mk(self.flat_grads)

    def views_of(self, buf: torch.Tensor, p: torch.Tensor, off: int) -> torch.Tensor:
        """ View of another flat buffer of the same layout (optimizer moments) with `p`'s shape and physical element order. """
        n = p.numel()
        if p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous() and not p.is_contiguous():
            k, c, r, s = p.shape
            return buf[off:off + n].view(k, r, s, c).permute(0, 3, 1, 2)
        return buf[off:off + n].view(p.shape)

    def grad_view(self, p: torch.Tensor) -> torch.Tensor:
        return self._grad_views[id(p)]

    def _wire_layers(self):
        """ Tell every fused layer where its parameter gradients live. """
        self.layers: List[FusedLayer] = [m for m in self.model.modules() if isinstance(m, FusedLayer)]
        wired = set()
        for layer in self.layers:
            targets = {}
            op, bn, gn = layer._op, layer._bn, layer._gn
            for key, prm in (('weight', op.weight), ('bias', op.bias), ('bn_w', getattr(bn, 'weight', None)), ('bn_b', getattr(bn, 'bias', None)),
                             ('gn_w', getattr(gn, 'weight', None)), ('gn_b', getattr(gn, 'bias', None))):
                if prm is not None and id(prm) in self._grad_views:
                    targets[key] = self._grad_views[id(prm)]
                    wired.add(id(prm))
            layer._grad_out = targets
        # parameters of modules the library does not own get their gradients from autograd's in-place accumulation: these need zeroing
        self.foreign_params = [p for p in self.params if id(p) not in wired]

    def restore_grad_views(self):
        """ `.grad` may have been set to None by a foreign `zero_grad(set_to_none=True)`: point it back into the flat buffer. """
        for p in self.params:
            if p.grad is not self._grad_views[id(p)]:
                p.grad = self._grad_views[id(p)]

    def reset_gradients(self, memset: bool = True):
        """ What `zero_grad()` means for the flat buffer: `.grad`s point into it again, the whole buffer is zero (ONE memset: parameters of layers that
        this step does not run, or of modules the library does not own, must not keep last step's values) and every fused layer writes — not
        accumulates — its gradients on its next backward. `memset=False` when the caller has zeroed the buffer itself (captured step: the arena's
        `begin_step(extra_zero=[flat_grads])`). """
        self.restore_grad_views()
        if memset:
            if self.flat_grads.is_cuda:
                check(lib.dcv_fill_zero(_ptr(self.flat_grads), self.flat_grads.numel() * 4, _stream()), 'fill_zero(flat gradients)')
            else:
                self.flat_grads.zero_()
        for layer in self.layers:
            if layer._grad_out is not None:
                layer._grad_out['_written'] = False


def flatten_parameters(model: torch.nn.Module, bucket_bytes: int = 8 << 20) -> FlatParameters:
    flat = getattr(model, '_flat_parameters', None)
    if flat is None:
        flat = FlatParameters(model, bucket_bytes)
        object.__setattr__(model, '_flat_parameters', flat)
    return flat


class FlatAdamW(torch.optim.Optimizer):
    """ `torch.optim.AdamW` semantics (decoupled weight decay, bias correction, amsgrad=False) as ONE kernel over the flat buffers.

    Drop-in for the `opt` type handed to `train()`: `FlatAdamW(model.parameters(), lr=..., betas=..., eps=..., weight_decay=...)`.
    The parameters must have been flattened (`flatten_parameters(model)`; `train()` does it) — otherwise each parameter is stepped with
    its own launch of the same kernel. The learning rate is read from a device scalar refreshed from `param_groups[0]['lr']` before each
    step (so LR schedulers keep working and a captured CUDA graph sees new values); the step count lives on the device too. """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, grad_scale: float = 1.0):
        if amsgrad:
            raise NotImplementedError('deepcv_b200: FlatAdamW does not build amsgrad (the reference recipe sets `amsgrad: false`)')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))
        self.grad_scale = float(grad_scale)
        self._flat: Optional[FlatParameters] = None
        self._dev_state = None

    def attach(self, flat: FlatParameters) -> 'FlatAdamW':
        self._flat = flat
        return self

    def zero_grad(self, set_to_none: bool = True):
        """ Flat mode: one memset of the flat gradient buffer (none inside a captured step whose arena zeroes it up front); the fused layers then
        write their gradients straight into their slices on the next backward. """
        if self._flat is None:
            super().zero_grad(set_to_none=set_to_none)
        else:
            self._flat.reset_gradients(memset=not self._flat.zeroed_by_step)

    def state_dict(self):
        """ torch layout plus the device-side step counter (read back once here), so that bias correction resumes where it stopped. """
        sd = super().state_dict()
        sd['flat_step'] = int(self._dev_state['step'].item()) if self._dev_state is not None else 0
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        step = int(state_dict.pop('flat_step', 0))
        super().load_state_dict(state_dict)
        st = self._device_state(self.param_groups[0]['params'][0].device)
        st['step'].fill_(step)
        st['lr_value'] = None

    def per_parameter_state(self) -> Dict[str, Dict[str, Any]]:
        """ `torch.optim.AdamW`-style state per parameter ({'step', 'exp_avg', 'exp_avg_sq'} views of the flat moments, in `model.named_parameters()` order):
        what a checkpoint needs to interchange with the stock optimizer (the reference saves 'optimizer' in `to_save`, `ignite_training.py:319`). """
        if self._flat is None or 'flat' not in self.state:
            return {}
        flat, st = self._flat, self.state['flat']
        step = torch.tensor(float(self._dev_state['step'].item()) if self._dev_state is not None else 0.)
        names = {id(p): n for n, p in flat.model.named_parameters()}
        out = {}
        for p, off in zip(flat.params, flat.offsets):
            v1, v2 = flat.views_of(st['exp_avg'], p, off), flat.views_of(st['exp_avg_sq'], p, off)
            out[names[id(p)]] = dict(step=step.clone(), exp_avg=v1, exp_avg_sq=v2)
        return out

    def _device_state(self, device):
        if self._dev_state is None:
            self._dev_state = dict(lr=torch.zeros((), dtype=torch.float32, device=device), step=torch.zeros((), dtype=torch.int32, device=device),
                                   lr_host=torch.zeros((), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else torch.zeros(()), lr_value=None)
        return self._dev_state

    def set_lr_device(self):
        """ Host -> device refresh of the learning rate (outside any CUDA graph). """
        group = self.param_groups[0]
        p0 = group['params'][0]
        st = self._device_state(p0.device)
        if st['lr_value'] != group['lr']:
            st['lr_host'].fill_(float(group['lr']))
            st['lr'].copy_(st['lr_host'], non_blocking=True)
            st['lr_value'] = group['lr']

    @torch.no_grad()
    def step(self, closure: Optional[Callable] = None, refresh_lr: bool = True):
        loss = closure() if closure is not None else None
        if len(self.param_groups) != 1 and self._flat is not None:
            raise NotImplementedError('deepcv_b200: FlatAdamW over flat buffers supports a single parameter group')
        group = self.param_groups[0]
        st = self._device_state(group['params'][0].device)
        if refresh_lr:
            self.set_lr_device()
        b1, b2 = group['betas']
        stream = _stream()
        check(lib.dcv_counter_add(_ptr(st['step']), 1, stream), 'counter_add')
        if self._flat is not None:
            flat = self._flat
            state = self.state.setdefault('flat', {})
            if not state:
                state['exp_avg'] = torch.zeros_like(flat.flat_params)
                state['exp_avg_sq'] = torch.zeros_like(flat.flat_params)
            check(lib.dcv_adamw_flat(_ptr(flat.flat_params), _ptr(flat.flat_grads), _ptr(state['exp_avg']), _ptr(state['exp_avg_sq']), flat.numel, _ptr(st['lr']),
                                     b1, b2, group['eps'], group['weight_decay'], self.grad_scale, _ptr(st['step']), stream), 'adamw_flat')
            return loss
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                g = p.grad
                if g.stride() != p.stride():
                    g = g.contiguous(memory_format=torch.channels_last) if (p.dim() == 4 and not p.is_contiguous()) else g.contiguous()
                state = self.state[p]
                if not state:
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                check(lib.dcv_adamw_flat(_ptr(p), _ptr(g), _ptr(state['exp_avg']), _ptr(state['exp_avg_sq']), p.numel(), _ptr(st['lr']),
                                         b1, b2, group['eps'], group['weight_decay'], self.grad_scale, _ptr(st['step']), stream), 'adamw_flat')
        return loss


class PeerAllReduce:
    """ One-shot all-reduce of small gradient buckets over NVLink peer memory (`dcv_peer_allreduce_sum`, csrc/peer_allreduce.cu): every rank pushes its
    slice into every peer's RECEIVE area (symmetric memory allocated through `torch.distributed._symmetric_memory` — PyTorch as the peer-mapping plumbing;
    the kernel is ours), then adds the W slices in rank order (bit-identical on all ranks). Replaces the collective library for buckets up to `max_floats`
    of models whose flat gradient buffer is at most `max_numel` floats (the receive area is 2 x world x numel floats); everything else stays on NCCL.
    Construction is collective (rendezvous). Disabled with DCV_NO_PEER_ALLREDUCE=1 or when symmetric memory is unavailable (`try_create` then returns
    None and the reducer uses NCCL for every bucket — logged, not silent). """
    max_numel = 1 << 20

    @staticmethod
    def enabled(world_size: int, device) -> bool:
        return (1 < world_size <= 8 and torch.device(device).type == 'cuda' and os.environ.get('DCV_NO_PEER_ALLREDUCE') is None
                and dist.is_initialized() and dist.get_backend() == 'nccl')

    def __init__(self, flat_grads: torch.Tensor, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        dev = flat_grads.device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.grads = flat_grads
        self.stride = (flat_grads.numel() + 3) // 4 * 4
        if self.stride > self.max_numel:
            raise ValueError(f'gradient buffer of {flat_grads.numel()} floats exceeds the peer all-reduce limit of {self.max_numel}')
        self.recv = symm_mem.empty(2 * self.world * self.stride, dtype=torch.float32, device=dev)
        self.recv.zero_()
        self._h_recv = symm_mem.rendezvous(self.recv, group)
        self.flags = symm_mem.empty(int(lib.dcv_peer_flag_words()), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self._h_flags = symm_mem.rendezvous(self.flags, group)
        self.state = torch.zeros(2 * 16, dtype=torch.int32, device=dev)
        self.max_floats = int(lib.dcv_peer_max_floats())
        self.max_slots = 16
        # device arrays of the peers' pointers (kept alive by the handles)
        self._recv_dev = ctypes.c_void_p(int(self._h_recv.buffer_ptrs_dev))
        self._flags_dev = ctypes.c_void_p(int(self._h_flags.buffer_ptrs_dev))
        torch.cuda.synchronize(dev)
        dist.barrier(group)   # every rank's flag words are zero before the first kernel may write into them

    @classmethod
    def try_create(cls, flat_grads: torch.Tensor, group=None):
        try:
            return cls(flat_grads, group)
        except Exception as e:   # no peer access / symmetric memory on this system, or too large a model: the collective library serves every bucket
            import logging
            logging.warning(f'deepcv_b200: peer-memory all-reduce not used ({type(e).__name__}: {e}); gradient buckets go through NCCL')
            return None

    def serves(self, start: int, end: int, slot: int) -> bool:
        return 0 < end - start <= self.max_floats and slot < self.max_slots and start % 4 == 0

    def all_reduce(self, start: int, end: int, slot: int):
        check(lib.dcv_peer_allreduce_sum(_ptr(self.grads), self._recv_dev, self._flags_dev, self.rank, self.world, self.stride, start, end - start, slot, _ptr(self.state), _stream()),
              'peer_allreduce_sum')


class GradientBucketReducer:
    """ Data-parallel gradient averaging over the flat gradient buffer: one `all_reduce(SUM)` per bucket on a communication stream,
    launched as soon as backward has enqueued the bucket's last producer. "Enqueued" is reported by the fused layers themselves, at the END of
    their backward (`ops._backward_done` -> the `_done` entry of `FusedLayer._grad_out`): a module full-backward hook fires when the gradient
    w.r.t. the layer's INPUT is ready, which for a first layer whose input needs no gradient is before its weight-gradient kernels are enqueued.
    Buckets that hold parameters of modules the library does not own, or of layers that did not run this step, are reduced by `finish()`.
    The 1/world_size factor is applied by `finish()` unless an optimizer has taken it over (`FlatAdamW.grad_scale`, then `average_in_finish`
    is False). BatchNorm statistics are never exchanged (per-replica BN, reference `use_sync_batch_norm: False`). """

    def __init__(self, flat: FlatParameters, process_group=None, overlap: bool = True):
        self.flat, self.group, self.overlap = flat, process_group, overlap
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.is_cuda = flat.flat_grads.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat.flat_grads.device) if self.is_cuda else None
        self.side_streams: List[torch.cuda.Stream] = []
        self.average_in_finish = True
        self.peer: Optional[PeerAllReduce] = None   # set by DataParallelModel: small buckets then go through the one-shot peer-memory kernel
        foreign = {flat.bucket_of[id(p)] for p in flat.foreign_params}
        self._early_ok = [b not in foreign for b in range(len(flat.buckets))]
        self._layer_buckets: Dict[int, List[int]] = {}
        # fused layers that write into each bucket: a bucket is complete when each of them has finished its backward
        self._producers = [0] * len(flat.buckets)
        for layer in flat.layers:
            touched = sorted({flat.bucket_of[id(p)] for p in layer.parameters(recurse=True) if id(p) in flat.bucket_of})
            self._layer_buckets[id(layer)] = touched
            for b in touched:
                self._producers[b] += 1
            if layer._grad_out is not None and overlap and self.world_size > 1:
                layer._grad_out['_done'] = self._make_done(layer)
        self.begin_step()

    def _make_done(self, layer):
        def _done():
            if id(layer) in self._finished:     # a second backward through the same layer (shared weights): its bucket waits for `finish()`
                for b in self._layer_buckets[id(layer)]:
                    self._early_ok_step[b] = False
                return
            self._finished.add(id(layer))
            for b in self._layer_buckets[id(layer)]:
                self._pending[b] -= 1
                if self._pending[b] == 0 and self._early_ok_step[b] and not self._launched[b]:
                    self._launch(b)
        return _done

    @property
    def inline_peer(self) -> bool:
        """ A model whose whole gradient buffer fits ONE one-shot peer all-reduce (the default net: 68 KB, ~9 us) is reduced by a single kernel on the
        COMPUTE stream at the end of backward: no communication stream inside the captured step (a forked graph costs more in node scheduling than the
        overlap of a few microseconds of exchange could return). DCV_PEER_OVERLAP=1 keeps the bucketed, overlapped schedule. """
        return self.peer is not None and self.flat.flat_grads.numel() <= self.peer.max_floats and os.environ.get('DCV_PEER_OVERLAP') is None

    def _launch(self, b: int):
        if self.inline_peer:
            return   # `finish()` reduces the whole buffer with one kernel
        start, end = self.flat.buckets[b]
        chunk = self.flat.flat_grads[start:end]
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            for side in self.side_streams:   # weight-gradient kernels of other layers of the bucket may still be running there (ops.StepContext.side_stream)
                self.comm_stream.wait_stream(side)
            with torch.cuda.stream(self.comm_stream):
                if self.peer is not None and self.peer.serves(start, end, b):
                    self.peer.all_reduce(start, end, b)      # one kernel over NVLink peer memory (small buckets: latency)
                else:
                    dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        self._launched[b] = True

    def begin_step(self):
        self._pending = list(self._producers)
        self._launched = [False] * len(self.flat.buckets)
        self._early_ok_step = list(self._early_ok)
        self._finished = set()

    def finish(self):
        """ After `backward()`: reduce whatever was not launched early, make the compute stream wait for the results, average. """
        if self.world_size <= 1:
            return
        relaunch = [b for b in range(len(self.flat.buckets)) if self._launched[b] and not self._early_ok_step[b]]
        if relaunch:
            raise RuntimeError('deepcv_b200: a gradient bucket was all-reduced before a second backward pass wrote into it (weights shared between layers '
                               'that complete at different times); construct the data-parallel wrap with `overlap=False`')
        if self.inline_peer:
            self.peer.all_reduce(0, self.flat.flat_grads.numel(), 0)
        else:
            for b in range(len(self.flat.buckets)):
                if not self._launched[b]:
                    self._launch(b)
            if self.is_cuda:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        if self.average_in_finish:
            g = self.flat.flat_grads
            if self.is_cuda:
                check(lib.dcv_axpby(_ptr(g), None, _ptr(g), 1. / self.world_size, 0., g.numel(), 0, _stream()), 'axpby(average gradients)')
            else:
                g.mul_(1. / self.world_size)
        self.begin_step()
