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
from .flat_params import FlatAdamW, FlatParameters, GradientBucketReducer, flatten_parameters
from .hyperparams import HYPERPARAMS_T, to_hyperparameters

__all__ = ['MAIN_TRAINING_LOSS_NAME', 'Events', 'State', 'Engine', 'PiecewiseLinear', 'BackendConfig', 'CrossEntropyLoss', 'train', 'make_process_function',
           'GraphedTrainStep', 'DataParallelModel']

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

    def __init__(self, module: torch.nn.Module, process_group=None, bucket_bytes: int = 8 << 20, overlap: bool = True):
        super().__init__()
        self.module = module
        self.flat = flatten_parameters(module, bucket_bytes)
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if self.world_size > 1:
            dist.broadcast(self.flat.flat_params, src=0, group=process_group)
            for b in module.buffers():
                dist.broadcast(b, src=0, group=process_group)
        self.reducer = GradientBucketReducer(self.flat, process_group, overlap=overlap)

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


def make_process_function(hp, device, model: torch.nn.Module, losses: 'OrderedDict[str, Callable]', optimizer: torch.optim.Optimizer) -> Callable[[Engine, Any], Dict[str, float]]:
    """ The training step of the reference, line for line (:233-255). """
    def process_function(engine: Engine, batch) -> Dict[str, float]:
        x, *y = tuple(b.to(device, non_blocking=True) if isinstance(b, torch.Tensor) and b.device != device else b for b in batch)
        if len(y) == 1:
            y = y[0]
        model.train()
        y_pred = model(x)
        batch_losses = {n: loss(y_pred, y) for n, loss in losses.items()}
        optimizer.zero_grad()
        batch_losses[MAIN_TRAINING_LOSS_NAME].backward()
        if isinstance(model, DataParallelModel):
            model.finish_gradient_reduction()
        optimizer.step()
        return {n: loss.item() for n, loss in batch_losses.items()}
    return process_function


class GraphedTrainStep:
    """ The same step captured ONCE into a CUDA graph (forward, backward, bucket all-reduces, AdamW) and replayed per batch.

    `step(x, y)` copies the batch into static buffers (x may be the raw uint8 batch when `model` starts with `FusedPreprocess`; the per-sample
    flip / crop parameters are drawn on the host and copied in too), refreshes the device learning rate, replays the graph and returns the
    loss tensor (device scalar, no sync). Requires an optimizer whose `step` is capture-safe (`FlatAdamW`). """

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: FlatAdamW, example_x: torch.Tensor, example_y: torch.Tensor, warmup_iters: int = 3,
                 preprocess: Optional[torch.nn.Module] = None, use_accumulator_arena: bool = True):
        if not example_x.is_cuda:
            raise RuntimeError('deepcv_b200: GraphedTrainStep needs CUDA tensors')
        self.model, self.loss_fn, self.optimizer, self.preprocess = model, loss_fn, optimizer, preprocess
        self.static_x, self.static_y = example_x.clone(), example_y.clone()
        n = example_x.shape[0]
        self.static_flip = self.static_crop = None
        self._host_flip = self._host_crop = None
        if preprocess is not None:
            self.static_flip = torch.zeros(n, dtype=torch.uint8, device=example_x.device)
            self.static_crop = torch.full((n, 2), int(preprocess.pad), dtype=torch.int32, device=example_x.device)
            self._host_flip = torch.zeros(n, dtype=torch.uint8).pin_memory()
            self._host_crop = torch.full((n, 2), int(preprocess.pad), dtype=torch.int32).pin_memory()
        self.graph = torch.cuda.CUDAGraph()
        self.static_loss = None
        model.train()
        optimizer.set_lr_device()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # One zeroed arena per step for all atomically-filled accumulators + ONE memset of the flat gradient buffer, instead of a memset node in front of
        # every kernel that accumulates (ops.AccumulatorArena). The last warm-up step runs in counting mode to size the arena.
        self.use_arena = use_accumulator_arena and os.environ.get('DCV_NO_ARENA') is None
        self.arena = ops.AccumulatorArena() if self.use_arena else None   # owned by this object: the graph replays write into its buffer
        # one bf16 cast of the flat parameter buffer per step instead of one per convolution (ops.set_param_shadows); bf16 steps only
        flat_p = getattr(optimizer, '_flat', None)
        op_dtype = getattr(preprocess, 'dtype', None) if preprocess is not None else (example_x.dtype if example_x.is_floating_point() else None)
        self._shadow = None
        if flat_p is not None and op_dtype == torch.bfloat16 and os.environ.get('DCV_NO_SHADOW') is None:
            self._shadow = torch.empty(flat_p.flat_params.numel(), dtype=torch.bfloat16, device=example_x.device)
            ops.set_param_shadows([(flat_p.flat_params, self._shadow)])
        try:
            self._capture(side, warmup_iters, example_x, optimizer)
        finally:
            ops.set_param_shadows([])   # eager forwards outside the captured step cast per layer: they must never see one-step-old shadows

    def _capture(self, side, warmup_iters, example_x, optimizer):
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
                flat = getattr(optimizer, '_flat', None)
                self.arena.begin_step(extra_zero=[flat.flat_grads] if flat is not None else [])
            try:
                self.static_loss = self._eager_step()
            finally:
                if self.use_arena:
                    self.arena.end_step()

    def _eager_step(self) -> torch.Tensor:
        if self._shadow is not None:
            ops.refresh_param_shadows()
        x = self.static_x
        if self.preprocess is not None:
            x = self.preprocess(x, flip=self.static_flip, crop_yx=self.static_crop)
        y_pred = self.model(x)
        loss = self.loss_fn(y_pred, self.static_y)
        self.optimizer.zero_grad()
        loss.backward()
        if isinstance(self.model, DataParallelModel):
            self.model.finish_gradient_reduction()
        self.optimizer.step(refresh_lr=False)
        return loss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        if self.preprocess is not None:
            flip, crop = self.preprocess.draw(x.shape[0])
            if flip is not None:
                self._host_flip.copy_(flip)
                self._host_crop.copy_(crop)
                self.static_flip.copy_(self._host_flip, non_blocking=True)
                self.static_crop.copy_(self._host_crop, non_blocking=True)
        self.optimizer.set_lr_device()
        self.graph.replay()
        return self.static_loss


def train(hp: HYPERPARAMS_T, model: torch.nn.Module, losses, datasets: Dict[str, Dataset], opt: Type[torch.optim.Optimizer] = FlatAdamW, backend_conf: BackendConfig = None,
          loss_weights=None, metrics: Dict[str, Callable] = None, callbacks_handler=None, nni_compression_pruner=None) -> Tuple[Dict[str, float], State]:
    """ Training procedure (reference :178-370, without the bookkeeping handlers). `datasets` maps 'trainset' (and optionally 'validset' /
    'testset') to datasets yielding `(x, y)`; returns (last batch losses, engine state). """
    TRAINING_HP_DEFAULTS = {'optimizer_opts': ..., 'epochs': ..., 'batch_size': ..., 'scheduler': None, 'output_path': Path.cwd() / 'data/04_training/',
                            'log_output_dir_to_mlflow': True, 'validate_every_epochs': 1, 'save_every_iters': 1000, 'log_grads_every_iters': -1, 'log_progress_every_iters': 100,
                            'seed': None, 'prefetch_batches': True, 'resume_from': '', 'crash_iteration': -1, 'deterministic_cudnn': False, 'use_sync_batch_norm': False,
                            'num_workers': 0}
    backend_conf = backend_conf if backend_conf is not None else BackendConfig()
    hp, _ = to_hyperparameters(hp, TRAINING_HP_DEFAULTS, raise_if_missing=True)
    device = backend_conf.device
    if hp['seed'] is not None:
        torch.manual_seed(backend_conf.rank + hp['seed'])  # a different seed per worker (reference :208)

    trainset = datasets['trainset'] if isinstance(datasets, dict) else datasets[0]
    sampler = torch.utils.data.distributed.DistributedSampler(trainset) if (backend_conf.distributed and dist.is_initialized()) else None
    train_loader = DataLoader(trainset, batch_size=hp['batch_size'], shuffle=sampler is None, sampler=sampler, num_workers=hp['num_workers'], pin_memory=backend_conf.is_cuda, drop_last=True)

    model = model.to(device)
    model = _setup_distributed_training(device, backend_conf, model, use_sync_batch_norm=hp['use_sync_batch_norm'])
    losses = _setup_ignite_losses(losses, loss_weights=loss_weights, device=device)
    optimizer = opt(model.parameters(), **hp['optimizer_opts'])
    if isinstance(optimizer, FlatAdamW):
        flat = flatten_parameters(model.module if isinstance(model, DataParallelModel) else model)
        optimizer.attach(flat)
        if isinstance(model, DataParallelModel):
            optimizer.grad_scale = 1. / model.world_size
    scheduler = None
    if hp['scheduler'] is not None:
        args_to_eval = hp['scheduler']['eval_args'] if 'eval_args' in hp['scheduler'] else {}
        scheduler_kwargs = {n: eval(v, {'hp': hp, 'iterations': len(train_loader)}) if n in args_to_eval else v for n, v in hp['scheduler']['kwargs'].items()}
        scheduler = hp['scheduler']['type'](optimizer=optimizer, **scheduler_kwargs)

    trainer = Engine(make_process_function(hp, device, model, losses, optimizer))
    if scheduler is not None:
        trainer.add_event_handler(Events.ITERATION_STARTED, scheduler)
    if sampler is not None:
        trainer.add_event_handler(Events.EPOCH_STARTED, lambda engine: sampler.set_epoch(engine.state.epoch))
    if hp['crash_iteration'] is not None and hp['crash_iteration'] >= 0:
        @trainer.on(Events.ITERATION_STARTED)
        def _(engine):
            if engine.state.iteration == hp['crash_iteration']:
                raise Exception(f'STOP at iteration: {engine.state.iteration}')

    state = trainer.run(train_loader, max_epochs=hp['epochs'])
    return state.output, state
