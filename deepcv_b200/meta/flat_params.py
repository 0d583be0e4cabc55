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
        """ Gradients are overwritten (not accumulated) by the backward kernels: nothing to clear in flat mode. """
        if self._flat is None:
            super().zero_grad(set_to_none=set_to_none)
        else:
            self._flat.restore_grad_views()
            for p in self._flat.foreign_params:
                p.grad.zero_()

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


class GradientBucketReducer:
    """ Data-parallel gradient averaging over the flat gradient buffer: one `all_reduce(SUM)` per bucket on a communication stream,
    launched as soon as backward has enqueued the bucket's last producer; the 1/world_size factor is folded into the optimizer
    (`FlatAdamW.grad_scale`). BatchNorm statistics are never exchanged (per-replica BN, reference `use_sync_batch_norm: False`). """

    def __init__(self, flat: FlatParameters, process_group=None, overlap: bool = True):
        self.flat, self.group, self.overlap = flat, process_group, overlap
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.is_cuda = flat.flat_grads.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat.flat_grads.device) if self.is_cuda else None
        # how many fused layers produce gradients into each bucket
        self._producers = [0] * len(flat.buckets)
        self._layer_buckets: Dict[int, List[int]] = {}
        for layer in flat.layers:
            touched = sorted({flat.bucket_of[id(p)] for p in layer.parameters(recurse=True) if id(p) in flat.bucket_of})
            self._layer_buckets[id(layer)] = touched
            for b in touched:
                self._producers[b] += 1
            if layer._grad_out is not None:
                layer._grad_out = dict(layer._grad_out)
        self._pending = list(self._producers)
        self._hooks = []
        if overlap and self.world_size > 1:
            for layer in flat.layers:
                self._hooks.append(layer.register_full_backward_hook(self._make_hook(layer)))

    def _make_hook(self, layer):
        def _hook(module, grad_input, grad_output):
            for b in self._layer_buckets[id(layer)]:
                self._pending[b] -= 1
                if self._pending[b] == 0:
                    self._launch(b)
        return _hook

    def _launch(self, b: int):
        start, end = self.flat.buckets[b]
        chunk = self.flat.flat_grads[start:end]
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        self._launched[b] = True

    def begin_step(self):
        self._pending = list(self._producers)
        self._launched = [False] * len(self.flat.buckets)

    def finish(self):
        """ After `backward()`: reduce whatever was not launched by the hooks, then make the compute stream wait for the results. """
        if self.world_size <= 1:
            return
        if not hasattr(self, '_launched'):
            self.begin_step()
        for b in range(len(self.flat.buckets)):
            if not self._launched[b]:
                self._launch(b)
        if self.is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.begin_step()
