""" Layer building blocks of the DeepcvModule path — host-side mirror of `src/deepcv/meta/nn.py`.

Same public names and argument meaning as the reference (`layer` :519-554, `normalization_techniques` :448-516, `NormTechnique`
:433-441, `get_padding_from_kernel` :393-399, `get_gain_name` :585-605, `interpolate` :665-676, `get_out_features_shape`
:689-704, `get_model_capacity` :679-686, `is_conv` / `is_fully_connected`), but the modules they return run the sm_100a
kernels of `deepcv_b200.ops` instead of ATen/cuDNN. The parameter containers are the very `torch.nn` modules the reference
instantiates (Conv2d, BatchNorm2d, GroupNorm, Linear ...), kept as children in the reference's order, so `state_dict()`
keys and shapes interchange with the reference / the CPU oracle; only `forward` differs.

Shape inference (the reference runs a dummy CPU forward, nn.py:689-704) is done with tensors on the `meta` device: every
module here answers a `meta` input with an empty `meta` output of the right shape. That is not a compute fallback: CPU
tensors are rejected.
"""
import enum
import functools
import logging
import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Type, Union

import numpy as np
import torch

from .. import ops
from .._lib import ACT_NONE, ALGO_AUTO

__all__ = ['XAVIER_INIT_SUPPORTED_ACT_FN', 'NormTechnique', 'NORM_TECHNIQUES_MODULES', 'batch_norm_nd', 'instance_norm_nd', 'avg_pooling_nd', 'conv_nd', 'forward_call_convention_dec', 'is_torch_obj',
           'get_padding_from_kernel', 'normalization_techniques', 'normalization_techniques_impl', 'layer', 'FusedLayer', 'AvgPool2d', 'Flatten',
           'get_gain_name', 'interpolate', 'get_model_capacity', 'get_out_features_shape', 'is_conv', 'is_fully_connected', 'is_torch_obj', 'meta_like']

XAVIER_INIT_SUPPORTED_ACT_FN = {torch.nn.ReLU: 'relu', torch.nn.LeakyReLU: 'leaky_relu', torch.nn.Tanh: 'tanh', torch.nn.Sigmoid: 'sigmoid', torch.nn.Identity: 'linear'}


def is_torch_obj(v) -> bool:
    """ reference nn.py:706-709 as intended (the `'_module__'` typo made it always False): a single tensor, not a sequence of
    tensors. A raw block output with a pending normalisation (`ops.PendingAffine`) counts as one tensor. """
    return isinstance(v, (torch.Tensor, torch.Size, ops.PendingAffine, ops.PendingNorm, ops.PendingFlatten))


def forward_call_convention_dec(apply_parallel_forward: bool = False, refs_tensor_count_similar: bool = None, in_tensors_count_similar_to_refs: bool = None,
                                ignore_prev_subm_intput: bool = False, ignore_sub_refs: bool = False):
    """ Call convention of submodule forward functions (reference nn.py:130-194): the decorated `forward_fn` always sees a LIST of tensors — or, with
    `apply_parallel_forward`, ONE tensor per call: the function is then applied to each input tensor in turn, together with the i-th tensor of every
    referenced submodule output ("parallel apply": siamese / parallel branches sharing one submodule) — whether the previous submodule produced a
    single tensor or a sequence; references arrive through `referenced_submodules_out` as `List[List[Tensor]]` (`List[Tensor]` when applied in
    parallel); the result is flattened and a single output tensor is returned bare. `refs_tensor_count_similar` / `in_tensors_count_similar_to_refs`
    (default: `apply_parallel_forward`) require all references (and the input) to hold the same number of tensors; `ignore_prev_subm_intput` /
    `ignore_sub_refs` drop the respective operand. Works on plain functions and on `torch.nn.Module.forward` methods.

    One deliberate difference (defect ledger): the reference consumes its operand lists with `list.pop()` — from the END — so
    every parallel submodule would silently reverse the order of the branches; here the i-th output belongs to the i-th input. """
    refs_tensor_count_similar = apply_parallel_forward if refs_tensor_count_similar is None else refs_tensor_count_similar
    refs_tensor_count_similar = not ignore_sub_refs and refs_tensor_count_similar
    in_tensors_count_similar_to_refs = apply_parallel_forward if in_tensors_count_similar_to_refs is None else in_tensors_count_similar_to_refs
    in_tensors_count_similar_to_refs = not ignore_sub_refs and not ignore_prev_subm_intput and in_tensors_count_similar_to_refs

    def _decorator(forward_fn: Callable) -> Callable:
        @functools.wraps(forward_fn)
        def _forward_wraper(*call_args, referenced_submodules_out=None, **kwargs):
            bound = ()
            if call_args and isinstance(call_args[0], torch.nn.Module):   # decorating a method: `self` comes first
                bound, call_args = call_args[:1], call_args[1:]
            inputs, args = (call_args[0] if call_args else None), call_args[1:]
            if ignore_prev_subm_intput or inputs is None:
                inputs = []
            else:
                inputs = [inputs] if is_torch_obj(inputs) else list(inputs)
            if ignore_sub_refs:
                refs = []
            else:
                refs = [[t] if is_torch_obj(t) else list(t) for t in (referenced_submodules_out or [])]
                if refs:
                    if in_tensors_count_similar_to_refs and not all(len(r) == len(inputs) for r in refs):
                        raise ValueError(f'Error: When `in_tensors_count_similar_to_refs` is `True`, all referenced output tensor(s) should each have as many tensor(s) as input '
                                         f'tensor(s) from previous submodule: `len(referenced_submodules_out[i] == len(inputs) =={len(inputs)}` {chr(10)}'
                                         f'Got {len(inputs)} input tensor(s) and references of {[len(r) for r in refs]} tensor(s)')
                    elif (ignore_prev_subm_intput or not in_tensors_count_similar_to_refs) and refs_tensor_count_similar and not all(len(r) == len(refs[0]) for r in refs):
                        raise ValueError(f'Error: When `refs_tensor_count_similar` is `True`, all referenced output tensor(s) should each have as many tensor(s) as the first '
                                         f'tensor(s) ref: `len(referenced_submodules_out[0])={len(refs[0])}` (or should all have only a single tensor if `refs[0]` isnt a sequence) '
                                         f'{chr(10)}Got references of {[len(r) for r in refs]} tensor(s)')
            if not apply_parallel_forward:
                refs_kwarg = {'referenced_submodules_out': refs} if refs else {}
                outs = forward_fn(*bound, inputs, *args, **refs_kwarg, **kwargs)
                outs = [outs] if is_torch_obj(outs) else list(outs)
            else:
                outs = []
                loops_count = 1   # at least one call, even when the previous output is ignored
                if not ignore_prev_subm_intput and inputs:
                    loops_count = len(inputs)
                if refs:
                    loops_count = len(refs[0])
                for i in range(loops_count):
                    refs_kwarg = {'referenced_submodules_out': [r[i] for r in refs]} if refs else {}
                    rslt = forward_fn(*bound, inputs[i] if inputs else [], *args, **refs_kwarg, **kwargs)
                    outs.extend([rslt] if is_torch_obj(rslt) else rslt)
            return outs if len(outs) != 1 else outs[0]
        return _forward_wraper
    return _decorator


def is_conv(op_t: Union[torch.nn.Module, Type]) -> bool:
    """ reference nn.py `is_conv` (tested at :731-740): class name contains 'conv', for a type or an instance. """
    t = op_t if isinstance(op_t, type) else type(op_t)
    return issubclass(t, torch.nn.Module) and 'conv' in t.__name__.lower()


def is_fully_connected(op_t: Union[torch.nn.Module, Type]) -> bool:
    t = op_t if isinstance(op_t, type) else type(op_t)
    return issubclass(t, torch.nn.Linear)


def get_padding_from_kernel(kernel_size, warn_on_uneven_kernel: bool = False):
    """ floor((k - 1) / 2) per dim (reference nn.py:393-399). """
    is_sequence = isinstance(kernel_size, Sequence)
    sizes = list(kernel_size) if is_sequence else [kernel_size]
    padding = [max(0, math.floor((ks - 1.) / 2.)) for ks in sizes]
    if warn_on_uneven_kernel and any(v % 2 == 0 for v in sizes):
        logging.warning(f'Warning: `kernel_size={kernel_size}` has even size, which may result in inapropriate output tensor shape even with "{padding}" zero-padding')
    return padding if is_sequence else padding[0]


def _nd(types: Sequence[Type], dims: int, what: str):
    if not 1 <= dims <= len(types):
        raise ValueError(f'Error: {what} is only available for 1D to {len(types)}D features, got `dims={dims}`')
    return types[dims - 1]


def conv_nd(dims: int = 2, **kwargs):
    return _nd((torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Conv3d), dims, 'convolution')(**kwargs)


def avg_pooling_nd(dims: int = 2, **kwargs):
    """ reference nn.py:416. The 2-D case (the one on the hot path) is the kernel-backed `AvgPool2d` below. """
    return _nd((torch.nn.AvgPool1d, AvgPool2d, torch.nn.AvgPool3d), dims, 'average pooling')(**kwargs)


def batch_norm_nd(dims: int = 2, **kwargs):
    """ reference nn.py:418 """
    return _nd((torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d), dims, 'batch normalization')(**kwargs)


def instance_norm_nd(dims: int = 2, **kwargs):
    return _nd((torch.nn.InstanceNorm1d, torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d), dims, 'instance normalization')(**kwargs)


class NormTechnique(enum.Enum):
    """ reference nn.py:433-441 """
    BATCH_NORM = r'batch_norm'
    LAYER_NORM = r'layer_norm'
    INSTANCE_NORM = r'instance_norm'
    GROUP_NORM = r'group_norm'
    LOCAL_RESPONSE_NORM = r'local_response_norm'
    LAYER_NORM_WITH_MEAN_ONLY_BATCH_NORM = r'layer_nrm_and_mean_batch_nrm'


NORM_TECHNIQUES_MODULES = {NormTechnique.BATCH_NORM: batch_norm_nd, NormTechnique.LAYER_NORM: torch.nn.LayerNorm, NormTechnique.INSTANCE_NORM: instance_norm_nd,
                           NormTechnique.GROUP_NORM: torch.nn.GroupNorm, NormTechnique.LOCAL_RESPONSE_NORM: torch.nn.LocalResponseNorm}


def normalization_techniques_impl(norm_type, norm_kwargs, input_shape=None, supported_norm_ops=NORM_TECHNIQUES_MODULES) -> List[torch.nn.Module]:
    """ reference nn.py:448-502: instantiates the normalisation modules, filling feature counts from `input_shape` (C, *spatial). """
    if isinstance(norm_type, (NormTechnique, str)):
        norm_type, norm_kwargs = [norm_type], [norm_kwargs]
    norm_type = [NormTechnique(t) if isinstance(t, str) else t for t in norm_type]
    norm_kwargs = list(norm_kwargs)
    if len(set(norm_type)) != len(norm_type):
        raise ValueError(f'Error: Cant use the same normalization technique mutiple times at once (duplicates forbiden in `norm_type` argument; Got `norm_type(s)="{norm_type}"`')
    if len(norm_type) != len(norm_kwargs):
        raise TypeError('Error: `norm_type` and `norm_kwargs` must either both be a sequence of the same size or both only one normalization technique and one keyword args dict; '
                        f'Got `norm_type(s)="{norm_type}"` and `norm_kwargs="{norm_kwargs}"`')
    norm_ops = []
    for norm_t, kwargs in zip(norm_type, norm_kwargs):
        kwargs = dict(kwargs)
        if input_shape is not None:
            if norm_t in (NormTechnique.INSTANCE_NORM, NormTechnique.BATCH_NORM):
                kwargs['num_features'] = input_shape[0]
                if len(input_shape) > 1:
                    kwargs['dims'] = len(input_shape) - 1
            if norm_t == NormTechnique.LAYER_NORM:
                kwargs['normalized_shape'] = list(input_shape[1:])
            elif norm_t == NormTechnique.GROUP_NORM:
                kwargs['num_channels'] = input_shape[0]
        if norm_t not in supported_norm_ops:
            raise ValueError(f'Error: "{norm_t}" is an unkown or forbiden normalization technique: It isn\'t specified in `supported_norm_ops="{supported_norm_ops}"`')
        norm_ops.append(supported_norm_ops[norm_t](**kwargs))
    return norm_ops


def normalization_techniques(input_shape=None, batch_norm: Dict[str, Any] = None, layer_norm: Dict[str, Any] = None, instance_norm: Dict[str, Any] = None,
                             group_norm: Dict[str, Any] = None, layer_nrm_and_mean_batch_nrm: Dict[str, Any] = None) -> List[torch.nn.Module]:
    """ reference nn.py:505-516: fixed order BN -> LN -> IN -> GN, only the configured (non-empty kwargs) ones. """
    techniques = {NormTechnique.BATCH_NORM: batch_norm, NormTechnique.LAYER_NORM: layer_norm, NormTechnique.INSTANCE_NORM: instance_norm,
                  NormTechnique.GROUP_NORM: group_norm, NormTechnique.LAYER_NORM_WITH_MEAN_ONLY_BATCH_NORM: layer_nrm_and_mean_batch_nrm}
    norms = {t: args for t, args in techniques.items() if args is not None and len(args) > 0}
    if not norms:
        return []
    return normalization_techniques_impl(list(norms.keys()), list(norms.values()), input_shape=input_shape)


def meta_like(shape: Sequence[int], dtype: torch.dtype = torch.float32) -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device='meta')


def _pair(v) -> tuple:
    return (int(v), int(v)) if isinstance(v, (int, np.integer)) else tuple(int(i) for i in v)


class AvgPool2d(torch.nn.AvgPool2d):
    """ `torch.nn.AvgPool2d(kernel_size, stride)` (no padding, floor mode — the defaults the reference's YAML uses) on the pooling kernels. """

    def __init__(self, kernel_size, stride=None, padding=0, ceil_mode: bool = False, count_include_pad: bool = True, divisor_override=None):
        super().__init__(kernel_size, stride=stride, padding=padding, ceil_mode=ceil_mode, count_include_pad=count_include_pad, divisor_override=divisor_override)
        if _pair(padding) != (0, 0) or ceil_mode or divisor_override is not None:
            raise NotImplementedError('deepcv_b200: average pooling with padding / ceil_mode / divisor_override is not built for sm_100a (no PyTorch fallback on this path)')

    accepts_pending_affine = True   # a raw block output: the pending normalisation is applied inside the pooling kernel (pool(A*y + B) = A*pool(y) + B)

    @forward_call_convention_dec(apply_parallel_forward=True, ignore_sub_refs=True)   # reference submodule_creators.py:175
    def forward(self, x) -> torch.Tensor:
        k, s = _pair(self.kernel_size), _pair(self.stride if self.stride is not None else self.kernel_size)
        if isinstance(x, ops.PendingAffine):
            if k == s and k[0] == k[1] and x.shape[2] % k[0] == 0 and x.shape[3] % k[0] == 0:
                return ops.sc_affine_pool(x, k[0])
            x = ops.materialize(x)
        if x.device.type == 'meta':
            return meta_like((x.shape[0], x.shape[1], (x.shape[2] - k[0]) // s[0] + 1, (x.shape[3] - k[1]) // s[1] + 1), x.dtype)
        return ops.avg_pool2d(x, k, s)   # a `PendingNorm` (tensor-core block output) with 2x2 / stride-2 windows: normalise + pool in one pass


class Flatten(torch.nn.Flatten):
    """ `torch.nn.Flatten()` of the logical N x C x H x W tensor (features in (C, H, W) order), from NHWC memory. The architecture
    parser substitutes this class wherever a spec names `torch.nn.Flatten`. """

    can_defer_flatten = True   # `defer_flatten=True`: the caller promises a consumer that reads the image tensor through the Flatten index map itself

    @forward_call_convention_dec(apply_parallel_forward=True, ignore_sub_refs=True)
    def forward(self, x: torch.Tensor, defer_flatten: bool = False) -> torch.Tensor:
        if defer_flatten and x.device.type == 'cuda' and x.dim() == 4 and self.start_dim == 1 and self.end_dim in (-1, 3):
            return ops.PendingFlatten(ops.as_nhwc(x))
        if x.device.type == 'meta' or x.dim() != 4 or self.start_dim != 1 or self.end_dim not in (-1, 3):
            if x.device.type != 'meta' and x.dim() == 4:
                raise NotImplementedError('deepcv_b200: only `Flatten(start_dim=1, end_dim=-1)` of image tensors is built')
            return super().forward(x)
        return ops.flatten_nchw(x)


class FusedLayer(torch.nn.Sequential):
    """ What `layer()` returns: a `torch.nn.Sequential` holding the reference's modules in the reference's order (so parameter
    names match: `0.weight`, `2.running_mean`, ...), executed as ONE fused operator chain:

        conv2d:  conv + bias + activation + per-(n,c) statistics  ->  finalize  ->  one affine per (n,c) for BatchNorm∘GroupNorm
        linear:  GEMM + bias + activation

    Both block orders of the reference (nn.py:553) are built — post-activation `(?Dropout) - op - act - (?norms)` (convolution epilogue fusion) and
    pre-activation `(?Dropout) - (?norms) - act - op` (`ops.pre_norm_act` in front of a plain convolution) — with ReLU / LeakyReLU / Sigmoid / no
    activation, BatchNorm / GroupNorm / InstanceNorm and element-wise Dropout (Philox mask, `ops.dropout`); anything else raises
    `NotImplementedError` at construction (never a silent PyTorch fallback). """

    def __init__(self, *modules: torch.nn.Module, preactivation: bool = False):
        super().__init__(*modules)
        self.preactivation = bool(preactivation)
        self.algo = ALGO_AUTO
        self._grad_out: Optional[dict] = None  # filled by flat_params.FlatParameters when gradients live in flat buckets
        self._step_ctx: Optional[ops.StepContext] = None  # set by the driver of a captured step (accumulator arena, parameter shadow)
        self._plan()

    def _plan(self):
        mods = list(self)
        # plain (unregistered) references: the children keep their Sequential index names in state_dict()
        ref = lambda name, m: object.__setattr__(self, name, m)
        ref('_op', next((m for m in mods if hasattr(m, 'weight') and (is_conv(m) or is_fully_connected(m))), None))
        if self._op is None:
            raise ValueError(f'Error: Bad layer operation module argument, no convolution / linear op found in "{mods}"')
        if not isinstance(self._op, (torch.nn.Conv2d, torch.nn.Linear)) or isinstance(self._op, torch.nn.ConvTranspose2d):
            raise NotImplementedError(f'deepcv_b200: "{type(self._op).__name__}" layers are outside the sm_100a hot path (Conv2d and Linear are built)')
        if isinstance(self._op, torch.nn.Conv2d):
            if self._op.groups != 1 or self._op.padding_mode != 'zeros' or isinstance(self._op.padding, str):
                raise NotImplementedError('deepcv_b200: grouped convolutions / non-zero padding modes are not built')
            # Keep the weight parameter physically [K][R][S][C] (logically still OIHW: state_dict interchange is unaffected)
            self._op.weight.data = self._op.weight.data.contiguous(memory_format=torch.channels_last)
        drop = [m for m in mods if isinstance(m, torch.nn.modules.dropout._DropoutNd)]
        if len(drop) > 1 or any(not isinstance(m, torch.nn.Dropout) for m in drop):
            raise NotImplementedError(f'deepcv_b200: only one element-wise `torch.nn.Dropout` per layer is built (got {drop})')
        ref('_drop', drop[0] if drop else None)     # always the first op of the block (reference nn.py:553), on the block's input
        self._drop_state: Optional[ops.DropoutState] = None
        norm_types = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.GroupNorm, torch.nn.modules.instancenorm._InstanceNorm, torch.nn.LayerNorm, torch.nn.LocalResponseNorm)
        acts = [m for m in mods if m is not self._op and m not in drop and not isinstance(m, norm_types)]
        if len(acts) > 1:
            raise NotImplementedError(f'deepcv_b200: more than one activation in a layer: {acts}')
        self._act, self._slope = ops.activation_code(acts[0] if acts else None)
        ref('_bn', next((m for m in mods if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)), None))
        gn = [m for m in mods if isinstance(m, (torch.nn.GroupNorm, torch.nn.modules.instancenorm._InstanceNorm))]
        if len(gn) > 1 or any(isinstance(m, (torch.nn.LayerNorm, torch.nn.LocalResponseNorm)) for m in mods):
            raise NotImplementedError('deepcv_b200: LayerNorm / LocalResponseNorm / InstanceNorm+GroupNorm stacks are not built (BatchNorm, GroupNorm, InstanceNorm are)')
        ref('_gn', gn[0] if gn else None)
        if isinstance(self._gn, torch.nn.modules.instancenorm._InstanceNorm) and self._gn.track_running_stats:
            raise NotImplementedError('deepcv_b200: InstanceNorm with running statistics is not built')
        if self.preactivation and self._bn is not None and self._bn.num_features != (self._op.in_channels if isinstance(self._op, torch.nn.Conv2d) else self._op.in_features):
            raise ValueError(f'Error: pre-activation BatchNorm normalises the block input ({getattr(self._op, "in_channels", None)} channels), got num_features={self._bn.num_features}')
        if isinstance(self._op, torch.nn.Linear) and (self._bn is not None or self._gn is not None):
            raise NotImplementedError('deepcv_b200: normalisation after a fully connected layer is not built (the reference specs set `batch_norm: null` there)')

    def _norm_config(self) -> ops.NormConfig:
        bn, gn = self._bn, self._gn
        groups = 1
        if gn is not None:
            groups = gn.num_groups if isinstance(gn, torch.nn.GroupNorm) else gn.num_features
        return ops.NormConfig(use_bn=bn is not None, bn_eps=bn.eps if bn is not None else 1e-5, bn_momentum=bn.momentum if bn is not None else 0.1,
                              use_gn=gn is not None, gn_groups=groups, gn_eps=gn.eps if gn is not None else 1e-5)

    def _meta_forward(self, x: torch.Tensor) -> torch.Tensor:
        op = self._op
        if isinstance(op, torch.nn.Linear):
            return meta_like((x.shape[0], op.out_features), torch.float32)
        k, s, p, d = op.kernel_size, op.stride, op.padding, op.dilation
        oh = (x.shape[2] + 2 * p[0] - d[0] * (k[0] - 1) - 1) // s[0] + 1
        ow = (x.shape[3] + 2 * p[1] - d[1] * (k[1] - 1) - 1) // s[1] + 1
        return meta_like((x.shape[0], op.out_channels, oh, ow), x.dtype)

    accepts_pending_flatten = property(lambda self: isinstance(self._op, torch.nn.Linear) and not self.preactivation and (self._drop is None or self._drop.p == 0.))
    accepts_pending_affine = True   # few-channel convolutions normalise their input while loading it
    can_defer_affine = True         # ... and may hand their own raw output on (`defer_affine=True`: the caller promises a consumer that accepts it)

    def _few_channel_path(self, x) -> bool:
        op = self._op
        return (isinstance(op, torch.nn.Conv2d) and not self.preactivation and self.algo == ALGO_AUTO and x.device.type == 'cuda' and len(x.shape) == 4
                and ops.sc_conv_supported(tuple(x.shape), op.weight, op.stride, op.padding, op.dilation, x.dtype))

    def _norm_kwargs(self) -> dict:
        bn, gn = self._bn, self._gn
        return dict(bn_weight=bn.weight if bn is not None else None, bn_bias=bn.bias if bn is not None else None,
                    running_mean=bn.running_mean if bn is not None else None, running_var=bn.running_var if bn is not None else None,
                    num_batches_tracked=bn.num_batches_tracked if bn is not None else None,
                    gn_weight=gn.weight if gn is not None else None, gn_bias=gn.bias if gn is not None else None)

    @forward_call_convention_dec(apply_parallel_forward=True, ignore_sub_refs=True)   # reference submodule_creators.py:254: one layer shared by parallel branches
    def forward(self, x, defer_affine: bool = False):
        if not isinstance(x, (ops.PendingAffine, ops.PendingNorm)) and x.device.type == 'meta':
            return self._meta_forward(x)
        op, bn, gn, drop = self._op, self._bn, self._gn, self._drop
        if isinstance(x, ops.PendingFlatten) and not self.accepts_pending_flatten:
            x = ops.materialize(x)
        if drop is not None and drop.training and drop.p != 0.:   # reference nn.py:553: Dropout is the first op of both block orders
            x = ops.materialize(x)
            if self._drop_state is None or self._drop_state.counter.device != x.device:
                self._drop_state = ops.DropoutState(x.device)
            x = ops.dropout(x, float(drop.p), self._drop_state)
        training = bn.training if bn is not None else self.training
        if self.preactivation:   # `(?Dropout) - (?norms) - Act - Layer`: the normalisation and the activation act on the block input
            x = ops.materialize(x)
            if isinstance(op, torch.nn.Linear):
                if x.dim() != 2:
                    x = ops.flatten_nchw(x) if x.dim() == 4 else x.reshape(x.shape[0], -1)
                x = ops.activation(x.contiguous(), self._act, self._slope)
                return ops.linear_act(x, op.weight, op.bias, ACT_NONE, 0., grad_out=self._grad_out, step_ctx=self._step_ctx)
            has_norm = bn is not None or gn is not None
            if has_norm:
                x = ops.pre_norm_act(x, self._norm_config(), training=training, act=self._act, slope=self._slope, grad_out=self._grad_out, step_ctx=self._step_ctx,
                                     **self._norm_kwargs())
            else:
                x = ops.activation(ops.as_nhwc(x), self._act, self._slope)
            return ops.conv_block(x, op.weight, op.bias, op.stride, op.padding, op.dilation, ACT_NONE, 0., norm=None, training=training, algo=self.algo,
                                  grad_out=self._grad_out, step_ctx=self._step_ctx, notify=not has_norm)
        if isinstance(x, ops.PendingNorm):
            x = ops.materialize(x)
        few = self._few_channel_path(x)
        if isinstance(x, ops.PendingAffine) and not few:
            x = ops.materialize(x)
        if isinstance(op, torch.nn.Linear):
            return ops.linear_act(x, op.weight, op.bias, self._act, self._slope, grad_out=self._grad_out, step_ctx=self._step_ctx)
        if few:
            out = ops.sc_conv_block(x, op.weight, op.bias, op.padding, self._act, self._slope, norm=self._norm_config(), training=training,
                                    grad_out=self._grad_out, step_ctx=self._step_ctx, **self._norm_kwargs())
            return out if defer_affine else ops.materialize(out)
        return ops.conv_block(x, op.weight, op.bias, op.stride, op.padding, op.dilation, self._act, self._slope, norm=self._norm_config(), training=training,
                              algo=self.algo, grad_out=self._grad_out, step_ctx=self._step_ctx, defer_apply=defer_affine, **self._norm_kwargs())


def layer(layer_op: torch.nn.Module, act_fn: Optional[Type[torch.nn.Module]], dropout_prob: float = None, preactivation: bool = False,
          input_shape=None, **norms_kwargs: Dict[str, Any]) -> torch.nn.Module:
    """ reference nn.py:519-554. Post-activation order `(?Dropout) - Layer - Act - (?norms)`; pre-activation order
    `(?Dropout) - (?norms) - Act - Layer`. Returns a `torch.nn.Sequential` (here: its fused subclass). """
    if not hasattr(layer_op, 'weight'):
        raise ValueError(f'Error: Bad layer operation module argument, no `weight` attribute found in layer_op="{layer_op}"')
    norm_ops = []
    if any(norms_kwargs.values()):
        if not preactivation:
            if input_shape is None:
                raise ValueError('Error: `input_shape` is needed to size normalization ops applied after the layer op')
            # normalised tensor = the op's output (reference nn.py:545-548); its shape comes from a meta-device pass
            with torch.no_grad():
                probe = FusedLayer(layer_op) if isinstance(layer_op, (torch.nn.Conv2d, torch.nn.Linear)) else layer_op
                input_shape = tuple(probe(meta_like((1, *input_shape))).shape[1:])
        if batch := norms_kwargs.get('batch_norm'):
            if dropout_prob not in (None, 0., 0) and batch:
                logging.warning('Warning: Dropout used along with normalization technique(s), like BatchNorm, may be unrecommended')
        norm_ops = normalization_techniques(input_shape, **norms_kwargs)
    drop = torch.nn.Dropout(p=dropout_prob) if (dropout_prob is not None and dropout_prob != 0.) else None
    act = act_fn() if act_fn is not None else None
    mods = (drop, *norm_ops, act, layer_op) if preactivation else (drop, layer_op, act, *norm_ops)
    return FusedLayer(*(m for m in mods if m is not None), preactivation=preactivation)


def get_gain_name(act_fn: Type[torch.nn.Module], default: str = 'relu', supported_act_fns: Dict[Type[torch.nn.Module], str] = XAVIER_INIT_SUPPORTED_ACT_FN) -> str:
    """ reference nn.py:585-605 """
    if act_fn in supported_act_fns:
        return supported_act_fns[act_fn]
    logging.warning(f'Warning: Unsupported activation function "{act_fn}", defaulting to "{default}" xavier initialization.')
    return default


def interpolate(tensors, out_spatial_shape, scaling_mode: str = None, align_corners: bool = False):
    """ reference nn.py:665-676 (identity when the spatial shape already matches). 2-D bilinear only on this path. """
    out_spatial_shape = tuple(int(s) for s in out_spatial_shape)
    if scaling_mode is None:
        scaling_mode = {1: 'linear', 2: 'bilinear', 3: 'trilinear'}.get(len(out_spatial_shape), 'nearest')

    def _interpolate(x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape[-len(out_spatial_shape):]) == out_spatial_shape:
            return x
        if x.device.type == 'meta':
            return meta_like((*x.shape[:-len(out_spatial_shape)], *out_spatial_shape), x.dtype)
        if scaling_mode != 'bilinear' or x.dim() != 4:
            raise NotImplementedError(f'deepcv_b200: "{scaling_mode}" interpolation of {x.dim() - 2}-D features is not built (2-D bilinear is)')
        return ops.bilinear_resize(x, out_spatial_shape, align_corners=align_corners)
    return _interpolate(tensors) if isinstance(tensors, torch.Tensor) else [_interpolate(t) for t in tensors]


def get_model_capacity(model: Optional[torch.nn.Module]) -> int:
    """ reference nn.py:679-686 """
    if model is None:
        return 0
    return int(sum(np.prod(param.shape) for param in model.parameters(recurse=True)))


def get_out_features_shape(input_shape, module: torch.nn.Module, use_minibatches: bool = True):
    """ reference nn.py:689-704: output shape(s) of `module` for a dummy input — here a `meta` tensor, so no kernel runs.
    Like the reference it leaves the module in eval mode. """
    single = isinstance(input_shape[0], (int, np.integer))
    shapes = [input_shape] if single else list(input_shape)
    module.eval()
    with torch.no_grad():
        dummy = [meta_like((1, *s) if use_minibatches else s) for s in shapes]
        outputs = module(dummy[0] if single else dummy)
    if isinstance(outputs, torch.Tensor):
        return outputs.shape
    return {n: r.shape for n, r in outputs.items()} if isinstance(outputs, dict) else [r.shape for r in outputs]
