""" Training step and its data-parallel wrap — host-side mirror of the live part of `src/deepcv/meta/ignite_training.py`.

Kept from the reference: `BackendConfig` (:78-117), `train(hp, model, losses, datasets, opt, backend_conf, ...)` (:178-370) with its
`TRAINING_HP_DEFAULTS`, the `process_function(engine, batch)` step (:233-255: forward, losses, `zero_grad`, `backward` on `main_loss`,
`step`, `.item()`), the scheduler construction from `eval_args` (:224-231) and `_setup_distributed_training` (:373-390: one process per
GPU, per-replica BatchNorm). pytorch-ignite is not installed in this image, so a minimal `Engine` / `Events` / `State` and the
`PiecewiseLinear` parameter scheduler the YAML names (`parameters.yml:103-108`) are provided here with ignite's call signatures.
Bookkeeping handlers (checkpoints, TensorBoard, MLflow, NNI) are orchestration outside the hot path and are not rebuilt.

What differs underneath: gradients live in flat buckets written by the backward kernels and are all-reduced per bucket over NCCL,
overlapped with the rest of backward (`flat_params.GradientBucketReducer`), instead of `DistributedDataParallel`; and the whole step
(forward, backward, all-reduce, AdamW) can be captured once into a CUDA graph and replayed (`GraphedTrainStep`), because the default
CIFAR-10 network is launch-latency-bound, not bandwidth-bound.
"""
import ctypes
import enum
import logging
import multiprocessing
import os
from collections import OrderedDict, defaultdict
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Type, Union

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset

from .. import ops
from .._lib import check, lib
from .flat_params import FlatAdamW, FlatParameters, GradientBucketReducer, PeerAllReduce, flatten_parameters
from .hyperparams import HYPERPARAMS_T, to_hyperparameters

__all__ = ['MAIN_TRAINING_LOSS_NAME', 'Events', 'State', 'Engine', 'PiecewiseLinear', 'BackendConfig', 'CrossEntropyLoss', 'train', 'make_process_function',
           'make_eager_process_function', 'GraphedTrainStep', 'GraphedEvalStep', 'DeviceDataLoader', 'evaluate', 'DataParallelModel']

MAIN_TRAINING_LOSS_NAME = 'main_loss'


# ------------------------------------------------------------------------------------------------------------------------------
# Minimal ignite surface (ignite.engine.Engine / Events / State; ignite.contrib.handlers.PiecewiseLinear)

class Events(enum.Enum):
    STARTED = 'started'
    EPOCH_STARTED = 'epoch_started'
    ITERATION_STARTED = 'iteration_started'
    ITERATION_COMPLETED = 'iteration_completed'
    EPOCH_COMPLETED = 'epoch_completed'
    COMPLETED = 'completed'


class State:
    def __init__(self):
        self.iteration, self.epoch, self.max_epochs, self.epoch_length = 0, 0, None, None
        self.output, self.batch, self.metrics, self.dataloader = None, None, {}, None


class Engine:
    """ `ignite.engine.Engine`: runs `process_function(engine, batch)` over the data for `max_epochs`, firing `Events`. """

    def __init__(self, process_function: Callable[['Engine', Any], Any]):
        self._process_function = process_function
        self._handlers: Dict[Events, List[Tuple[Callable, tuple, dict]]] = defaultdict(list)
        self.state = State()
        self.should_terminate = False
        self._dataloader_iter = None

    def add_event_handler(self, event: Events, handler: Callable, *args, **kwargs):
        self._handlers[event].append((handler, args, kwargs))
        return handler

    def on(self, event: Events, *args, **kwargs):
        def _decorator(handler):
            self.add_event_handler(event, handler, *args, **kwargs)
            return handler
        return _decorator

    def _fire(self, event: Events):
        for handler, args, kwargs in self._handlers[event]:
            handler(self, *args, **kwargs)

    def terminate(self):
        self.should_terminate = True

    def run(self, data: Iterable, max_epochs: int = 1, epoch_length: Optional[int] = None) -> State:
        self.state.max_epochs, self.state.dataloader = max_epochs, data
        self.state.epoch_length = epoch_length if epoch_length is not None else (len(data) if hasattr(data, '__len__') else None)
        self._fire(Events.STARTED)
        while self.state.epoch < max_epochs and not self.should_terminate:
            self.state.epoch += 1
            self._fire(Events.EPOCH_STARTED)
            self._dataloader_iter = iter(data)
            for i, batch in enumerate(self._dataloader_iter):
                if self.should_terminate or (epoch_length is not None and i >= epoch_length):
                    break
                self.state.iteration += 1
                self.state.batch = batch
                self._fire(Events.ITERATION_STARTED)
                self.state.output = self._process_function(self, batch)
                self._fire(Events.ITERATION_COMPLETED)
            self._fire(Events.EPOCH_COMPLETED)
        self._fire(Events.COMPLETED)
        return self.state


class PiecewiseLinear:
    """ `ignite.contrib.handlers.PiecewiseLinear(optimizer, param_name, milestones_values)`: linear interpolation of an optimizer
    parameter between (event index, value) milestones; called as an `ITERATION_STARTED` handler. """

    def __init__(self, optimizer: torch.optim.Optimizer, param_name: str, milestones_values: Sequence[Tuple[int, float]], save_history: bool = False, param_group_index: Optional[int] = None):
        if not isinstance(milestones_values, Sequence) or len(milestones_values) < 1:
            raise ValueError(f'Argument milestones_values should be with at least one value, but given {milestones_values}')
        values, milestones = [], []
        for pair in milestones_values:
            if not isinstance(pair, Sequence) or len(pair) != 2:
                raise ValueError('Argument milestones_values should be a list of pairs (milestone, param_value)')
            if not isinstance(pair[0], int):
                raise ValueError(f'Value of a milestone should be integer, but given {type(pair[0])}')
            if len(milestones) > 0 and pair[0] < milestones[-1]:
                raise ValueError(f'Milestones should be increasing integers, but given {pair[0]} is smaller than the previous milestone {milestones[-1]}')
            milestones.append(pair[0])
            values.append(float(pair[1]))
        self.optimizer, self.param_name, self.values, self.milestones = optimizer, param_name, values, milestones
        self.param_group_index = param_group_index
        self.event_index, self._index = 0, 0

    def get_param(self) -> float:
        if self.event_index <= self.milestones[0]:
            return self.values[0]
        if self.event_index >= self.milestones[-1]:
            return self.values[-1]
        while self._index + 1 < len(self.milestones) and self.event_index >= self.milestones[self._index + 1]:
            self._index += 1
        while self._index > 0 and self.event_index < self.milestones[self._index]:
            self._index -= 1
        start, end = self.milestones[self._index], self.milestones[self._index + 1]
        v0, v1 = self.values[self._index], self.values[self._index + 1]
        return v0 + (v1 - v0) * (self.event_index - start) / (end - start)

    def __call__(self, engine: Optional[Engine] = None, name: Optional[str] = None):
        value = self.get_param()
        groups = self.optimizer.param_groups if self.param_group_index is None else [self.optimizer.param_groups[self.param_group_index]]
        for group in groups:
            group[self.param_name] = value
        self.event_index += 1


# ------------------------------------------------------------------------------------------------------------------------------

class BackendConfig:
    """ Device and distributed configuration (reference :78-117). `distributed` is True iff both `dist_backend` and `dist_url` are given. """

    def __init__(self, device_or_id: Union[None, str, int, torch.device] = None, dist_backend: str = None, dist_url: str = None):
        if device_or_id is None:
            self.device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        elif isinstance(device_or_id, (str, torch.device)):
            self.device = torch.device(device_or_id) if isinstance(device_or_id, str) else device_or_id
        else:
            self.device = torch.device('cuda', int(device_or_id))
        self.is_cpu = self.device.type == 'cpu'
        self.is_cuda = self.device.type == 'cuda'
        self.ncpu = multiprocessing.cpu_count()
        self.dist_backend = dist_backend
        self.dist_url = dist_url
        self.local_rank = getattr(self.device, 'index', None)
        self.ngpus_current_node = torch.cuda.device_count()
        self.rank, self.nnodes, self.gpus_world_size = 0, 1, 1
        if self.distributed and dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank()
            self.gpus_world_size = dist.get_world_size()
            self.nnodes = max(1, dist.get_world_size() // max(1, self.ngpus_current_node))

    @property
    def distributed(self) -> bool:
        return self.dist_backend is not None and self.dist_backend != '' and self.dist_url is not None and self.dist_url != ''

    def __str__(self) -> str:
        if self.is_cpu:
            return f'single-node-cpu-{self.ncpu}'
        if self.distributed:
            return f'distributed-{self.nnodes}avg_nodes-{self.gpus_world_size}gpus_world_size-{self.ngpus_current_node}current-node-gpus(rank={self.rank})'
        return f'single-node-{self.ngpus_current_node}-available-gpus'

    __repr__ = __str__


class CrossEntropyLoss(torch.nn.Module):
    """ `torch.nn.CrossEntropyLoss()` (mean reduction, class-index targets; the loss `classification/image.py:70` uses) on the library's
    fused log-softmax + NLL kernel. """

    def forward(self, y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return ops.cross_entropy(y_pred, y)


def _setup_ignite_losses(losses, loss_weights=None, device=None) -> 'OrderedDict[str, Callable]':
    """ Single loss / sequence / mapping of losses -> mapping whose first entry is `main_loss` (reference :138-176, single-term case:
    the weighted multi-term mean is host-side bookkeeping and uses ordinary tensor arithmetic). """
    if callable(losses) and not isinstance(losses, (dict, list, tuple)):
        return OrderedDict([(MAIN_TRAINING_LOSS_NAME, losses)])
    if isinstance(losses, dict):
        named = OrderedDict(losses)
    else:
        named = OrderedDict((f'loss_{i}', l) for i, l in enumerate(losses))
    if len(named) == 1:
        return OrderedDict([(MAIN_TRAINING_LOSS_NAME, next(iter(named.values())))])
    if loss_weights is None:
        weights = {n: 1. for n in named}
    else:
        weights = dict(loss_weights) if isinstance(loss_weights, dict) else {n: w for n, w in zip(named, loss_weights)}
    total_w = sum(weights.values())

    def _main(y_pred, y):
        return sum(weights[n] * l(y_pred, y) for n, l in named.items()) / total_w
    return OrderedDict([(MAIN_TRAINING_LOSS_NAME, _main), *named.items()])


class DataParallelModel(torch.nn.Module):
    """ What `_setup_distributed_training` returns in place of `DistributedDataParallel(model)`: same forward, gradients averaged across
    ranks by bucketed NCCL all-reduce overlapped with backward. Parameters are broadcast from rank 0 at construction (as DDP does);
    BatchNorm statistics stay per replica. """

    def __init__(self, module: torch.nn.Module, process_group=None, bucket_bytes: Optional[int] = None, overlap: bool = True):
        super().__init__()
        self.module = module
        if bucket_bytes is None:
            # 8 MB buckets (DDP's scale) for large models; a small model still gets ~4 buckets, so that all but the last (the first layers' few KB)
            # reduce while backward runs and only one small-message latency is exposed (the default 68 KB net was ONE bucket launched after the
            # last weight gradient). DCV_BUCKET_BYTES overrides.
            total = sum(p.numel() for p in module.parameters()) * 4
            bucket_bytes = int(os.environ.get('DCV_BUCKET_BYTES', 0)) or min(8 << 20, max(8 << 10, total // 4))
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.flat = flatten_parameters(module, bucket_bytes)
        if self.world_size > 1:
            dist.broadcast(self.flat.flat_params, src=0, group=process_group)
            for b in module.buffers():
                dist.broadcast(b, src=0, group=process_group)
        self.reducer = GradientBucketReducer(self.flat, process_group, overlap=overlap)
        if PeerAllReduce.enabled(self.world_size, self.flat.flat_grads.device) and self.flat.flat_grads.numel() <= PeerAllReduce.max_numel:
            self.reducer.peer = PeerAllReduce.try_create(self.flat.flat_grads, process_group)

    def forward(self, *args, **kwargs):
        self.reducer.begin_step()
        return self.module(*args, **kwargs)

    def finish_gradient_reduction(self):
        self.reducer.finish()


def _setup_distributed_training(device, backend_conf: BackendConfig, model: torch.nn.Module, batch_shape=None, use_sync_batch_norm: bool = False) -> torch.nn.Module:
    """ reference :373-390. One process per GPU; `use_sync_batch_norm` is not built (north star: per-replica BatchNorm). """
    if backend_conf.distributed:
        if not dist.is_initialized():
            dist.init_process_group(backend_conf.dist_backend, init_method=backend_conf.dist_url)
        if backend_conf.is_cuda:
            torch.cuda.set_device(backend_conf.device)
        if use_sync_batch_norm:
            raise NotImplementedError('deepcv_b200: SyncBatchNorm is not built (BatchNorm statistics are per replica on this path, the reference default)')
        return DataParallelModel(model)
    return model


def _fused_layers(model: torch.nn.Module):
    from .nn import FusedLayer
    return [m for m in model.modules() if isinstance(m, FusedLayer)]


def _move_batch(batch, device):
    x, *y = tuple(b.to(device, non_blocking=True) if isinstance(b, torch.Tensor) and b.device != device else b for b in batch)
    return x, (y[0] if len(y) == 1 else y)


def make_eager_process_function(hp, device, model: torch.nn.Module, losses: 'OrderedDict[str, Callable]', optimizer: torch.optim.Optimizer,
                                preprocess: Optional[torch.nn.Module] = None) -> Callable[[Engine, Any], Dict[str, float]]:
    """ The training step of the reference, line for line (:233-255), one kernel launch after the other. """
    inner = model.module if isinstance(model, DataParallelModel) else model
    flat = getattr(inner, '_flat_parameters', None)

    def process_function(engine: Engine, batch) -> Dict[str, float]:
        x, y = _move_batch(batch, device)
        model.train()
        if preprocess is not None:
            preprocess.train()
            x = preprocess(x)
        y_pred = model(x)
        batch_losses = {n: loss(y_pred, y) for n, loss in losses.items()}
        optimizer.zero_grad()
        if flat is not None and not isinstance(optimizer, FlatAdamW):
            flat.reset_gradients()   # a stock optimizer's zero_grad(set_to_none=True) dropped the `.grad` views into the flat buffer
        batch_losses[MAIN_TRAINING_LOSS_NAME].backward()
        if isinstance(model, DataParallelModel):
            model.finish_gradient_reduction()
        optimizer.step()
        return {n: loss.item() for n, loss in batch_losses.items()}
    return process_function


def make_process_function(hp, device, model: torch.nn.Module, losses: 'OrderedDict[str, Callable]', optimizer: torch.optim.Optimizer,
                          preprocess: Optional[torch.nn.Module] = None, cuda_graph: Optional[bool] = None) -> Callable[[Engine, Any], Dict[str, float]]:
    """ `process_function(engine, batch) -> {loss name: float}` of the reference (:233-255). With a `FlatAdamW` over flat buffers on a CUDA device the
    step is captured into a CUDA graph the first time a batch shape is seen (`GraphedTrainStep`, model / optimizer / BatchNorm state rewound after the
    warm-up steps) and replayed afterwards — the default network's step is launch-bound: 2.8 ms eager, 0.6 ms replayed. `cuda_graph=False` (or any
    other optimizer) keeps the eager step. At most 4 distinct batch shapes are captured (`drop_last=True` loaders have one). """
    capturable = isinstance(optimizer, FlatAdamW) and optimizer._flat is not None and torch.device(device).type == 'cuda' and os.environ.get('DCV_NO_GRAPH') is None
    if cuda_graph is None:
        cuda_graph = capturable
    if cuda_graph and not capturable:
        raise RuntimeError('deepcv_b200: a CUDA-graph training step needs a CUDA device and a `FlatAdamW` attached to the flattened parameters of the model')
    eager = make_eager_process_function(hp, device, model, losses, optimizer, preprocess)
    if not cuda_graph:
        return eager
    runners: Dict[Any, GraphedTrainStep] = {}

    dev = torch.device(device)

    def _on_device(t: torch.Tensor) -> bool:
        return t.is_cuda and (dev.index is None or t.device.index == dev.index)

    def _key(x, y):
        return (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype) if (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)) else None

    def process_function(engine: Engine, batch) -> Dict[str, float]:
        x, *y = batch
        y = y[0] if len(y) == 1 else y
        if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)):
            return eager(engine, batch)
        key = _key(x, y)
        runner = runners.get(key)
        if runner is None:
            if len(runners) >= 4:
                raise RuntimeError(f'deepcv_b200: more than 4 distinct batch shapes in one training run (got {key}); use a `drop_last=True` loader or `cuda_graph=False`')
            # batches that already live on the device (DeviceDataLoader yields the same buffers every time) become the graph's static inputs: no copy
            # per step; host batches (pinned, from a DataLoader) are copied straight into the static buffers by `step`
            adopt = _on_device(x) and _on_device(y)
            runner = runners[key] = GraphedTrainStep(model, losses, optimizer, x.to(dev, non_blocking=True), y.to(dev, non_blocking=True), preprocess=preprocess,
                                                     restore_state=True, adopt_inputs=adopt)
        runner.step(x, y)
        # the step is launched; whatever the host still has to do for the NEXT step hides under its device time: the copy of the next batch (a
        # prefetching loader: reference meta/data/datasets.py:76-115) and the draw of its augmentation parameters. Then the one sync of the step.
        it = getattr(engine, '_dataloader_iter', None)
        prefetch = getattr(it, 'prefetch', None)
        if prefetch is not None:
            prefetch()
        runner.prepare_next()
        pending = getattr(it, 'pending', None)
        if pending is not None:   # the next batch is on its way to the device: its copy into the graph's static inputs is enqueued now, behind this step
            nxt = pending()
            if nxt is not None and isinstance(nxt[0], (tuple, list)) and len(nxt[0]) == 2 and runners.get(_key(*nxt[0])) is runner:
                runner.stage_next(nxt[0][0], nxt[0][1], nxt[1])
        return {n: float(v) for n, v in zip(runner.static_losses.keys(), runner.read_losses())}
    process_function.runners = runners
    return process_function


class GraphedTrainStep:
    """ The same step captured ONCE into a CUDA graph (forward, backward, bucket all-reduces, AdamW) and replayed per batch.

    `step(x, y)` copies the batch into static buffers (x may be the raw uint8 batch when `preprocess` is a `FusedPreprocess`; the per-sample
    flip / crop parameters are drawn on the host and copied in too), refreshes the device learning rate, replays the graph and returns the
    main loss tensor (device scalar, no sync); `static_losses` holds every loss term. Requires an optimizer whose `step` is capture-safe (`FlatAdamW`).
    `restore_state`: parameters, BatchNorm statistics and optimizer state are put back (in place) to what they were before the warm-up / capture
    steps. `adopt_inputs`: `example_x` / `example_y` themselves become the static buffers (a loader that always yields the same tensors then needs
    no copy). All per-step resources (accumulator arena, bf16 parameter shadow) live in a `StepContext` owned by this object. """

    def __init__(self, model: torch.nn.Module, loss_fn: Union[Callable, 'OrderedDict[str, Callable]'], optimizer: FlatAdamW, example_x: torch.Tensor, example_y: torch.Tensor,
                 warmup_iters: int = 3, preprocess: Optional[torch.nn.Module] = None, use_accumulator_arena: bool = True, restore_state: bool = False, adopt_inputs: bool = False):
        if not example_x.is_cuda:
            raise RuntimeError('deepcv_b200: GraphedTrainStep needs CUDA tensors')
        self.model, self.optimizer, self.preprocess = model, optimizer, preprocess
        self.losses = loss_fn if isinstance(loss_fn, dict) else OrderedDict([(MAIN_TRAINING_LOSS_NAME, loss_fn)])
        self.static_x, self.static_y = (example_x, example_y) if adopt_inputs else (example_x.clone(), example_y.clone())
        n = example_x.shape[0]
        self.static_flip = self.static_crop = None
        self._host_flip = self._host_crop = None
        self._host_slot, self._staged = 0, None
        if preprocess is not None:
            self.static_flip = torch.zeros(n, dtype=torch.uint8, device=example_x.device)
            self.static_crop = torch.full((n, 2), int(preprocess.pad), dtype=torch.int32, device=example_x.device)
            self._host_flip = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._host_crop = [torch.full((n, 2), int(preprocess.pad), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.graph = torch.cuda.CUDAGraph()
        self.static_losses: 'OrderedDict[str, torch.Tensor]' = OrderedDict()
        self.static_loss = None
        self._drawn = None
        self._loss_pack = torch.empty(len(self.losses), dtype=torch.float32, device=example_x.device)
        self._loss_host = torch.empty(len(self.losses), dtype=torch.float32).pin_memory()
        self._loss_event = torch.cuda.Event(external=True)   # an event-record NODE of the captured graph, waited for by the host
        self._loss_stream = torch.cuda.Stream(device=example_x.device) if os.environ.get('DCV_LOSS_INLINE') is None else None
        inner = model.module if isinstance(model, DataParallelModel) else model
        self.flat: Optional[FlatParameters] = getattr(optimizer, '_flat', None) or getattr(inner, '_flat_parameters', None)
        if isinstance(optimizer, FlatAdamW) and optimizer._flat is None and self.flat is not None:
            optimizer.attach(self.flat)   # a flattened model's fused layers write into the flat buffer: the optimizer must step that buffer
        if self.flat is not None and not isinstance(optimizer, FlatAdamW):
            raise RuntimeError('deepcv_b200: the model\'s parameters live in flat buffers (flatten_parameters / DataParallelModel); a captured step needs `FlatAdamW`')
        snapshot = self._snapshot(inner) if restore_state else None
        model.train()
        if preprocess is not None:
            preprocess.train()
        optimizer.set_lr_device()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # One zeroed arena per step for all atomically-filled accumulators + ONE memset of the flat gradient buffer, instead of a memset node in front of
        # every kernel that accumulates (ops.AccumulatorArena). The last warm-up step runs in counting mode to size the arena.
        self.use_arena = use_accumulator_arena and os.environ.get('DCV_NO_ARENA') is None
        self.ctx = ops.StepContext(ops.AccumulatorArena() if self.use_arena else None)   # owned by this object: the graph replays write into its buffers
        self.arena = self.ctx.arena
        # the weight-gradient kernels of the tensor-core convolutions run on a second stream, concurrently with the HBM-bound normalisation backward of the
        # layer before (ops._ConvBlock.backward); only with the flat gradient buffer (gradients written in place, nothing handed back to autograd)
        if self.flat is not None and os.environ.get('DCV_SIDE_WGRAD') == '1':   # opt-in (see ops._SIDE_WGRAD)
            self.ctx.side_stream = torch.cuda.Stream(device=example_x.device)
            reducer = getattr(model, 'reducer', None)
            if reducer is not None:
                reducer.side_streams = [self.ctx.side_stream]
        # one bf16 cast of the flat parameter buffer per step instead of one per convolution; bf16 steps only
        op_dtype = getattr(preprocess, 'dtype', None) if preprocess is not None else (example_x.dtype if example_x.is_floating_point() else None)
        self._shadow = None
        if self.flat is not None and op_dtype == torch.bfloat16 and os.environ.get('DCV_NO_SHADOW') is None:
            self._shadow = torch.empty(self.flat.flat_params.numel(), dtype=torch.bfloat16, device=example_x.device)
            self.ctx.shadows = [(self.flat.flat_params, self._shadow)]
            # ... and one launch for all the transposed + flipped weight operands of the data-gradient convolutions (tensor-core layers: 64-channel multiples)
            convs = [l._op.weight for l in _fused_layers(inner) if isinstance(l._op, torch.nn.Conv2d) and l._op.in_channels % 64 == 0 and l._op.out_channels % 16 == 0
                     and l._op.stride == (1, 1)]
            if convs and os.environ.get('DCV_NO_BATCHED_PACK') is None:
                self.ctx.plan_transposed_weights(self.flat.flat_params, [w.detach() for w in convs], torch.bfloat16)
        layers = _fused_layers(inner)
        previous = [l._step_ctx for l in layers]
        for l in layers:
            l._step_ctx = self.ctx
        try:
            self._capture(side, max(1, warmup_iters), example_x)
        finally:
            # eager forwards outside the captured step allocate / cast per layer: they must never see this step's arena or one-step-old shadows
            for l, prev in zip(layers, previous):
                l._step_ctx = prev
            if self.flat is not None:
                self.flat.zeroed_by_step = False
        if snapshot is not None:
            self._restore(inner, snapshot)

    # ---- state rewind (in place: the captured graph holds the addresses)
    def _snapshot(self, inner):
        opt = self.optimizer
        st = opt.state.get('flat') if isinstance(opt, FlatAdamW) else None
        return dict(params=[p.detach().clone() for p in inner.parameters()], buffers=[b.detach().clone() for b in inner.buffers()],
                    moments=None if not st else (st['exp_avg'].clone(), st['exp_avg_sq'].clone()),
                    step=int(opt._dev_state['step'].item()) if getattr(opt, '_dev_state', None) is not None else 0,
                    rng=None if self.preprocess is None else self.preprocess.generator.get_state())

    def _restore(self, inner, snap):
        with torch.no_grad():
            for p, v in zip(inner.parameters(), snap['params']):
                p.copy_(v)
            for b, v in zip(inner.buffers(), snap['buffers']):
                b.copy_(v)
            opt = self.optimizer
            st = opt.state.get('flat') if isinstance(opt, FlatAdamW) else None
            if st:
                if snap['moments'] is None:
                    st['exp_avg'].zero_(), st['exp_avg_sq'].zero_()
                else:
                    st['exp_avg'].copy_(snap['moments'][0]), st['exp_avg_sq'].copy_(snap['moments'][1])
            if getattr(opt, '_dev_state', None) is not None:
                opt._dev_state['step'].fill_(snap['step'])
        if snap['rng'] is not None:
            self.preprocess.generator.set_state(snap['rng'])

    def _capture(self, side, warmup_iters, example_x):
        with torch.cuda.stream(side):
            for i in range(warmup_iters):
                if self.use_arena and i == warmup_iters - 1:
                    self.arena.measure()
                self._eager_step()
            if self.use_arena:
                self.arena.end_measure(example_x.device)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            if self.use_arena:
                self.arena.begin_step(extra_zero=[self.flat.flat_grads] if self.flat is not None else [])
                if self.flat is not None:
                    self.flat.zeroed_by_step = True
            try:
                self._eager_step()
            finally:
                if self.use_arena:
                    self.arena.end_step()
                if self.flat is not None:
                    self.flat.zeroed_by_step = False

    def _eager_step(self) -> torch.Tensor:
        # Drop the previous iteration's losses first: they keep its autograd graph alive, and with it the parameters' AccumulateGrad nodes, which remember
        # the stream they were created on (the warm-up side stream) — reused inside the capture, the engine would make the capturing stream wait for
        # that uncaptured stream at the end of backward ("dependency created on uncaptured work in another stream").
        self.static_losses.clear()
        self.static_loss = None
        self.ctx.refresh_shadows()
        x = self.static_x
        if self.preprocess is not None:
            x = self.preprocess(x, flip=self.static_flip, crop_yx=self.static_crop)
        y_pred = self.model(x)
        for name, fn in self.losses.items():
            self.static_losses[name] = fn(y_pred, self.static_y)
        loss = self.static_loss = self.static_losses[MAIN_TRAINING_LOSS_NAME]
        # The loss terms leave for the host HERE, before backward: a device -> pinned-host copy and an EXTERNAL event inside the captured step. The host's
        # one synchronisation per step (`read_losses`, the reference's `.item()`) then returns while backward is still running, and everything the
        # host does until the next replay — engine bookkeeping, next batch, graph launch — hides under it: the device never waits for the host.
        vals = [v.detach() for v in self.static_losses.values()]
        if len(vals) > 1:
            torch.stack(vals, out=self._loss_pack)
        main = torch.cuda.current_stream()
        if self._loss_stream is not None:   # the copy and the event are a side branch of the captured graph: backward does not queue behind a memcpy node
            self._loss_stream.wait_stream(main)
        with torch.cuda.stream(self._loss_stream if self._loss_stream is not None else main):
            self._loss_host.copy_(vals[0].reshape(1) if len(vals) == 1 else self._loss_pack, non_blocking=True)
            self._loss_event.record()
        self.optimizer.zero_grad()
        torch.autograd.backward(loss, grad_tensors=[ops.unit_grad(loss.device)])   # = loss.backward() without the ones_like fill kernel / the multiplication by one
        self.ctx.join_side()
        if self._loss_stream is not None:
            torch.cuda.current_stream().wait_stream(self._loss_stream)
        if isinstance(self.model, DataParallelModel):
            self.model.finish_gradient_reduction()
        self.optimizer.step(refresh_lr=False)
        return loss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self._staged != (x.data_ptr(), y.data_ptr()):   # else: `stage_next` already copied this very batch into the static buffers
            if x.data_ptr() != self.static_x.data_ptr():
                self.static_x.copy_(x, non_blocking=True)
            if y.data_ptr() != self.static_y.data_ptr():
                self.static_y.copy_(y, non_blocking=True)
        self._staged = None
        if self.preprocess is not None:
            if not self._drawn:
                self.prepare_next()
            self._drawn = None
        self.optimizer.set_lr_device()
        self.graph.replay()
        return self.static_loss

    def prepare_next(self) -> None:
        """ Draws the NEXT step's per-sample augmentation parameters (the draw order of the generator is unchanged) and enqueues their host -> device copy
        into the static buffers. Called right after a step has been launched, all of it hides under that step's device time: the copies are ordered
        behind the running step on the stream (it still reads the buffers), and the host has nothing left to do between that step's loss read and
        the next replay. Two pinned staging buffers alternate: the previous copy may not have executed yet (one synchronisation per step). """
        if self.preprocess is None or self._drawn:
            return
        self.preprocess.train()   # the captured step is a TRAINING step (reference :246 `model.train()`): an evaluation pass in between must not freeze the draws
        flip, crop = self.preprocess.draw(self.static_x.shape[0])
        if flip is not None:
            self._host_slot ^= 1
            host_flip, host_crop = self._host_flip[self._host_slot], self._host_crop[self._host_slot]
            host_flip.copy_(flip)
            host_crop.copy_(crop)
            self.static_flip.copy_(host_flip, non_blocking=True)
            self.static_crop.copy_(host_crop, non_blocking=True)
        self._drawn = 'staged'

    def stage_next(self, x: torch.Tensor, y: torch.Tensor, ready: Optional[torch.cuda.Event] = None) -> bool:
        """ Copies the NEXT batch (device tensors, e.g. the batch a prefetching loader has in flight; `ready`: the event of its arrival) into the static
        input buffers now — behind the step that is running — so that `step(x, y)` with these tensors only replays the graph. False (and nothing
        done) when the tensors do not match the captured shapes. """
        if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) and x.is_cuda and y.is_cuda and x.shape == self.static_x.shape and x.dtype == self.static_x.dtype
                and y.shape == self.static_y.shape and y.dtype == self.static_y.dtype):
            return False
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.static_y.data_ptr():
            self.static_y.copy_(y, non_blocking=True)
        self._staged = (x.data_ptr(), y.data_ptr())
        return True

    def read_losses(self) -> List[float]:
        """ The loss terms of the step just replayed, as floats (the step's only synchronisation, the reference's `.item()`): waits for the forward
        pass of that step, not for its backward / optimizer kernels. """
        self._loss_event.synchronize()   # recorded inside the replayed graph right after the losses' device -> host copy (see `_eager_step`)
        return self._loss_host.tolist()


# ------------------------------------------------------------------------------------------------------------------------------
# Input pipeline (reference: DataLoader + `dataloader_prefetch_batches`, ignite_training.py:211-218, meta/data/datasets.py:76-115)

class DeviceDataLoader:
    """ The whole dataset resident in HBM (uint8 H x W x C images when the recipe is `FusedPreprocess`, else whatever the per-sample transforms
    produce) and batches assembled on the device: per epoch ONE host -> device copy of the sampler's index list, per batch ONE launch of the
    row-gather kernel (`dcv_gather_rows`) per tensor, writing into buffers that never move — a captured training step adopts them as its static
    inputs, so no per-sample copy, no collate, no pinned staging, no per-batch copy. Sampling follows `torch.utils.data`: a host-side
    `randperm` (seeded generator) when `shuffle`, `DistributedSampler` sharding (pad to a multiple of the world size, rank-strided) when
    `world_size > 1`, `drop_last` semantics as in `DataLoader`. """

    def __init__(self, dataset: Dataset, batch_size: int, device, shuffle: bool = False, drop_last: bool = False, seed: int = 0, rank: int = 0, world_size: int = 1,
                 max_resident_bytes: int = 32 << 30):
        self.device, self.batch_size, self.shuffle, self.drop_last = torch.device(device), int(batch_size), bool(shuffle), bool(drop_last)
        self.rank, self.world_size, self.seed, self.epoch = int(rank), int(world_size), int(seed), 0
        first = dataset[0]
        if not isinstance(first, tuple):
            raise TypeError('DeviceDataLoader expects a dataset of (input, target, ...) tuples')
        n = len(dataset)
        columns = [[] for _ in first]
        nbytes = sum(torch.as_tensor(v).numel() * torch.as_tensor(v).element_size() for v in first) * n
        if nbytes > max_resident_bytes:
            raise MemoryError(f'deepcv_b200: dataset of {nbytes / 2**30:.1f} GiB exceeds `max_resident_bytes` ({max_resident_bytes / 2**30:.1f} GiB)')
        for i in range(n):
            for col, v in zip(columns, dataset[i]):
                col.append(torch.as_tensor(v))
        self.columns = [torch.stack(col).contiguous().to(self.device) for col in columns]      # one H2D copy per column, once
        self.n = n
        self._batch_bufs = [torch.empty((self.batch_size, *c.shape[1:]), dtype=c.dtype, device=self.device) for c in self.columns]
        self._err = torch.zeros((), dtype=torch.int32, device=self.device)
        self._idx_dev = None

    def set_epoch(self, epoch: int):
        self.epoch = int(epoch)

    def _indices(self) -> torch.Tensor:
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(self.n, generator=g)
        else:
            idx = torch.arange(self.n)
        if self.world_size > 1:   # DistributedSampler: pad by wrapping around, then rank-strided shards
            total = (self.n + self.world_size - 1) // self.world_size * self.world_size
            if total > self.n:
                idx = torch.cat([idx, idx[: total - self.n]])
            idx = idx[self.rank:total:self.world_size]
        return idx.contiguous()

    def __len__(self) -> int:
        per_rank = (self.n + self.world_size - 1) // self.world_size
        return per_rank // self.batch_size if self.drop_last else (per_rank + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = self._indices()
        self._idx_dev = idx.pin_memory().to(self.device, non_blocking=True)
        st = ops._stream()
        for b in range(len(self)):
            lo = b * self.batch_size
            cnt = min(self.batch_size, idx.numel() - lo)
            out = []
            for col, buf in zip(self.columns, self._batch_bufs):
                row_bytes = col[0].numel() * col.element_size()
                dst = buf if cnt == self.batch_size else buf[:cnt]
                check(lib.dcv_gather_rows(ops._ptr(col), ctypes.c_void_p(self._idx_dev.data_ptr() + lo * 8), ops._ptr(dst), cnt, self.n, row_bytes, ops._ptr(self._err), st), 'gather_rows')
                out.append(dst)
            yield tuple(out)
        if int(self._err.item()) != 0:
            raise IndexError('DeviceDataLoader: a sampler index was outside the dataset')


def _find_fused_preprocess(dataset) -> Optional[torch.nn.Module]:
    """ The `FusedPreprocess` of a `PreprocessedDataset` built from a fused recipe (its per-sample call keeps images uint8; the arithmetic runs per
    batch on the device, at the top of the step). """
    from .data.preprocess import FusedPreprocess
    tf = getattr(dataset, '_img_transform', None)
    for t in ([tf] if isinstance(tf, FusedPreprocess) else list(getattr(tf, 'transforms', []) or [])):
        if isinstance(t, FusedPreprocess):
            return t
    return None


# ------------------------------------------------------------------------------------------------------------------------------
# Evaluation (reference: `create_supervised_evaluator` runs every `validate_every_epochs` epochs and at the end, :288-307)

class GraphedEvalStep:
    """ Forward in eval mode (BatchNorm on running statistics, no autograd tape) + classification metrics, captured once per batch shape and replayed:
    the accumulators `acc3` = (sum of cross entropies, correct predictions, rows) stay on the device until `result()`. """

    def __init__(self, model: torch.nn.Module, example_x: torch.Tensor, example_y: torch.Tensor, preprocess: Optional[torch.nn.Module] = None, adopt_inputs: bool = False):
        self.model, self.preprocess = model, preprocess
        self.static_x, self.static_y = (example_x, example_y) if adopt_inputs else (example_x.clone(), example_y.clone())
        self.acc3 = torch.zeros(3, dtype=torch.float32, device=example_x.device)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self._eager()
        self.acc3.zero_()

    def _eager(self):
        was_training = self.model.training
        pre_was_training = self.preprocess.training if self.preprocess is not None else False
        self.model.eval()
        if self.preprocess is not None:
            self.preprocess.eval()
        try:
            with torch.no_grad():
                x = self.static_x if self.preprocess is None else self.preprocess(self.static_x)
                logits = self.model(x)
                logits = ops._cast_raw(logits.contiguous(), torch.float32)
                check(lib.dcv_classification_metrics(ops._ptr(logits), ops._ptr(self.static_y), ops._ptr(self.acc3), logits.shape[0], logits.shape[1], ops._stream()), 'classification_metrics')
        finally:
            self.model.train(was_training)
            if self.preprocess is not None:
                self.preprocess.train(pre_was_training)   # a preprocess left in eval mode would stop drawing flips / crop offsets for the training steps

    def step(self, x, y):
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.static_y.data_ptr():
            self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()

    def result(self, reset: bool = True) -> Dict[str, float]:
        loss_sum, correct, rows = self.acc3.tolist()
        if reset:
            self.acc3.zero_()
        return {'loss': loss_sum / max(rows, 1.), 'accuracy': correct / max(rows, 1.), 'samples': int(rows)}


def evaluate(model: torch.nn.Module, data: Iterable, device, preprocess: Optional[torch.nn.Module] = None, runners: Optional[dict] = None) -> Dict[str, float]:
    """ Loss (cross entropy) and accuracy of a classifier over `data` (batches of (x, y)), the metrics the reference's evaluators compute. """
    runners = {} if runners is None else runners
    inner = model.module if isinstance(model, DataParallelModel) else model
    total = {'loss': 0., 'accuracy': 0., 'samples': 0}
    for batch in data:
        x, y = _move_batch(batch, device)
        key = (tuple(x.shape), x.dtype)
        runner = runners.get(key)
        if runner is None:
            runner = runners[key] = GraphedEvalStep(inner, x, y, preprocess=preprocess)
        runner.step(x, y)
    for runner in runners.values():
        r = runner.result()
        total['loss'] += r['loss'] * r['samples']; total['accuracy'] += r['accuracy'] * r['samples']; total['samples'] += r['samples']
    n = max(total['samples'], 1)
    return {'loss': total['loss'] / n, 'accuracy': total['accuracy'] / n, 'samples': total['samples']}


def train(hp: HYPERPARAMS_T, model: torch.nn.Module, losses, datasets: Dict[str, Dataset], opt: Type[torch.optim.Optimizer] = FlatAdamW, backend_conf: BackendConfig = None,
          loss_weights=None, metrics: Dict[str, Callable] = None, callbacks_handler=None, nni_compression_pruner=None) -> Tuple[Dict[str, float], State]:
    """ Training procedure (reference :178-370, without the bookkeeping handlers). `datasets` maps 'trainset' (and optionally 'validset' /
    'testset') to datasets yielding `(x, y)`; returns (validation metrics — the last batch losses when there is no validation set —, engine state).
    On a CUDA device the datasets are made resident in HBM and batched by `DeviceDataLoader` (hp `device_resident_dataset`, default True; falls back
    to `torch.utils.data.DataLoader` when a dataset exceeds hp `max_resident_bytes`), the step is CUDA-graph replayed when `opt` is `FlatAdamW`, and
    the validation set is evaluated every `validate_every_epochs` epochs and at the end with batches 32 times the training batch (:214). """
    TRAINING_HP_DEFAULTS = {'optimizer_opts': ..., 'epochs': ..., 'batch_size': ..., 'scheduler': None, 'output_path': Path.cwd() / 'data/04_training/',
                            'log_output_dir_to_mlflow': True, 'validate_every_epochs': 1, 'save_every_iters': 1000, 'log_grads_every_iters': -1, 'log_progress_every_iters': 100,
                            'seed': None, 'prefetch_batches': True, 'resume_from': '', 'crash_iteration': -1, 'deterministic_cudnn': False, 'use_sync_batch_norm': False,
                            'num_workers': 0, 'device_resident_dataset': True, 'max_resident_bytes': 32 << 30, 'cuda_graph': None}
    backend_conf = backend_conf if backend_conf is not None else BackendConfig()
    hp, _ = to_hyperparameters(hp, TRAINING_HP_DEFAULTS, raise_if_missing=True)
    device = backend_conf.device
    seed = 0 if hp['seed'] is None else int(hp['seed'])
    if hp['seed'] is not None:
        torch.manual_seed(backend_conf.rank + hp['seed'])  # a different seed per worker (reference :208)

    named = dict(datasets) if isinstance(datasets, dict) else dict(zip(('trainset', 'validset', 'testset'), datasets))
    trainset = named['trainset']
    preprocess = _find_fused_preprocess(trainset)
    if preprocess is not None:
        preprocess = preprocess.to(device)
    distributed = backend_conf.distributed and dist.is_available()
    model = model.to(device)
    model = _setup_distributed_training(device, backend_conf, model, use_sync_batch_norm=hp['use_sync_batch_norm'])
    world, rank = (dist.get_world_size(), dist.get_rank()) if (distributed and dist.is_initialized()) else (1, 0)

    def _loader(ds, batch_size, is_train):
        if backend_conf.is_cuda and hp['device_resident_dataset']:
            try:
                return DeviceDataLoader(ds, batch_size, device, shuffle=is_train, drop_last=is_train, seed=seed, rank=rank if is_train else 0, world_size=world if is_train else 1,
                                        max_resident_bytes=hp['max_resident_bytes'])
            except MemoryError as e:
                logging.warning(f'{e}; falling back to torch.utils.data.DataLoader')
        sampler = torch.utils.data.distributed.DistributedSampler(ds) if (is_train and world > 1) else None
        return DataLoader(ds, batch_size=batch_size, shuffle=is_train and sampler is None, sampler=sampler, num_workers=hp['num_workers'], pin_memory=backend_conf.is_cuda, drop_last=is_train)
    train_loader = _loader(trainset, hp['batch_size'], True)
    if hp['prefetch_batches'] and backend_conf.is_cuda and isinstance(train_loader, DataLoader):   # reference :217-218 (host-resident datasets only: a DeviceDataLoader has nothing to copy)
        from .data.datasets import dataloader_prefetch_batches
        train_loader = dataloader_prefetch_batches(train_loader, device)
    eval_sets = {n: ds for n, ds in named.items() if n != 'trainset' and ds is not None}

    losses = _setup_ignite_losses(losses, loss_weights=loss_weights, device=device)
    optimizer = opt(model.parameters(), **hp['optimizer_opts'])
    inner = model.module if isinstance(model, DataParallelModel) else model
    if isinstance(optimizer, FlatAdamW):
        optimizer.attach(flatten_parameters(inner))
        if isinstance(model, DataParallelModel):   # the optimizer kernel applies 1/world_size: no separate averaging pass
            optimizer.grad_scale = 1. / model.world_size
            model.reducer.average_in_finish = False
    scheduler = None
    if hp['scheduler'] is not None:
        args_to_eval = hp['scheduler']['eval_args'] if 'eval_args' in hp['scheduler'] else {}
        scheduler_kwargs = {n: eval(v, {'hp': hp, 'iterations': len(train_loader)}) if n in args_to_eval else v for n, v in hp['scheduler']['kwargs'].items()}
        scheduler = hp['scheduler']['type'](optimizer=optimizer, **scheduler_kwargs)

    trainer = Engine(make_process_function(hp, device, model, losses, optimizer, preprocess=preprocess, cuda_graph=hp['cuda_graph']))
    if scheduler is not None:
        trainer.add_event_handler(Events.ITERATION_STARTED, scheduler)
    if hasattr(train_loader, 'set_epoch'):
        trainer.add_event_handler(Events.EPOCH_STARTED, lambda engine: train_loader.set_epoch(engine.state.epoch))
    elif getattr(train_loader, 'sampler', None) is not None and hasattr(train_loader.sampler, 'set_epoch'):
        trainer.add_event_handler(Events.EPOCH_STARTED, lambda engine: train_loader.sampler.set_epoch(engine.state.epoch))
    if hp['crash_iteration'] is not None and hp['crash_iteration'] >= 0:
        @trainer.on(Events.ITERATION_STARTED)
        def _(engine):
            if engine.state.iteration == hp['crash_iteration']:
                raise Exception(f'STOP at iteration: {engine.state.iteration}')

    eval_metrics: Dict[str, float] = {}
    if eval_sets and backend_conf.is_cuda:
        name0 = 'validset' if 'validset' in eval_sets else next(iter(eval_sets))
        eval_loader = _loader(eval_sets[name0], hp['batch_size'] * 32, False)
        eval_runners: dict = {}

        def _run_validation(engine: Engine):
            every = hp['validate_every_epochs']
            if engine.state.epoch == engine.state.max_epochs or (every and engine.state.epoch % every == 0):
                m = evaluate(model, eval_loader, device, preprocess=preprocess, runners=eval_runners)
                eval_metrics.update({f'valid_{k}': v for k, v in m.items()})
                engine.state.metrics.update(eval_metrics)
        trainer.add_event_handler(Events.EPOCH_COMPLETED, _run_validation)

    state = trainer.run(train_loader, max_epochs=hp['epochs'])
    return (dict(eval_metrics) if eval_metrics else state.output), state
