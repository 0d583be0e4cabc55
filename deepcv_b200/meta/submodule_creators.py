""" Submodule-creator registry of the DeepcvModule path — host-side mirror of `src/deepcv/meta/submodule_creators.py`.

This is the plugin API the sm_100a operators sit behind: `BASIC_SUBMODULE_CREATORS` (:38-40),
`submodule_creator_dec` (:133-160), `ForwardCallbackSubmodule` (:85-128), the reduction functions (:43-65) and the built-in
creators `conv2d`, `linear` / `fully_connected`, `average_pooling` (+ the `avg_pooling` spelling the YAML uses),
`residual_link`, `dense_link`, `reduce`, `_new_branch_from_tensor` (:163-332). Creator names, arguments and error behaviour follow
the reference (with its defect ledger applied); what the created modules *execute* are the kernels of
`deepcv_b200.ops`. Creators the reference registers for other paths (conv1d/3d, transposed convolutions, HRNet blocks,
coordinate maps) are outside the hot path: naming them raises `NotImplementedError` instead of silently running PyTorch.
"""
import inspect
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence, Set, Type, Union

import numpy as np
import torch

from .. import ops
from . import nn as deepcv_nn

__all__ = ['BASIC_SUBMODULE_CREATORS', 'TENSOR_REDUCTION_FNS', 'get_reduction_fn', 'ForwardCallbackSubmodule', 'submodule_creator_dec', 'avg_pooling_creator',
           'reduction_subm_creator', 'select_tensor_creator', 'parse_slice', 'new_branch_creator', 'add_nn_layer_creator', 'add_residual_dense_link_creator', 'FROM', 'FROM_NAS_INPUT_CHOICE', 'NEW_BRANCH_FROM_TENSOR']

# Token values of `deepcv.meta.nn_spec.yaml_tokens` used by creators (kept here as strings to avoid an import cycle)
FROM, FROM_NAS_INPUT_CHOICE, NEW_BRANCH_FROM_TENSOR = '_from', '_from_nas_input_choice', '_new_branch_from_tensor'
NL = '\n'

BASIC_SUBMODULE_CREATORS: Dict[str, Callable[..., torch.nn.Module]] = {}


def get_reduction_fn(reduction: str) -> Callable:
    """ 'mean' / 'sum' (elementwise over the tensor list), 'concat' (channel dim) or 'none' (list unchanged) — the intended
    semantics of reference :43-65. A single tensor passes through. """
    if reduction not in ('mean', 'sum', 'concat', 'none'):
        raise ValueError(f'Error: Invalid "{reduction}" reduction function name. Valid reduction functions are "mean", "sum", "concat" and "none".')

    def _reduction_fn(tensors, keep_dim: bool = False, out: Optional[torch.Tensor] = None):
        if reduction == 'none' or isinstance(tensors, torch.Tensor):
            return tensors
        tensors = list(tensors)
        if tensors[0].device.type == 'meta':
            shape = list(tensors[0].shape)
            if reduction == 'concat':
                shape[1] = sum(t.shape[1] for t in tensors)
            return deepcv_nn.meta_like(shape, tensors[0].dtype)
        return ops.link_reduce(tensors, reduction)
    return _reduction_fn


TENSOR_REDUCTION_FNS = {reduction: get_reduction_fn(reduction) for reduction in ['mean', 'sum', 'concat', 'none']}


class ForwardCallbackSubmodule(torch.nn.Module):
    """ Module defined by a forward callback (reference :85-128). `DeepcvModule` wires `referenced_submodules` from the spec's
    `_from` entry and calls `subm(x, referenced_submodules_out=OrderedDict[name -> tensor])`. """

    def __init__(self, forward_callback: Callable):
        super().__init__()
        self.forward_callback = forward_callback
        self.takes_tensor_references = 'referenced_submodules_out' in inspect.signature(forward_callback).parameters
        self.mutable_input_choice = None   # NNI NAS InputChoice: not part of this path
        self.referenced_submodules: Optional[List[str]] = None

    def forward(self, tensors, referenced_submodules_out: 'OrderedDict[str, torch.Tensor]' = None):
        if referenced_submodules_out is not None and self.referenced_submodules is not None and self.takes_tensor_references:
            refs = [v for n, v in referenced_submodules_out.items() if n in self.referenced_submodules]   # each a tensor or a list of tensors (parallel branches)
            return self.forward_callback(tensors, referenced_submodules_out=refs)
        if referenced_submodules_out or self.referenced_submodules or self.takes_tensor_references:
            raise ValueError(f'Error: Uncoherent usage of output tensor references: (Did you provided `{FROM}` to a submodule which doesnt support tensor references?){NL}'
                             f'Got `referenced_submodules_out="{referenced_submodules_out}"` while `self.referenced_submodules="{self.referenced_submodules}"` '
                             f'and `self.takes_tensor_references="{self.takes_tensor_references}"`')
        return self.forward_callback(tensors)


def submodule_creator_dec(name: str, submodule_creators: Dict[str, Callable] = BASIC_SUBMODULE_CREATORS, allowed_subm_params_keys: Set[str] = None,
                          required_subm_params_keys: Set[str] = None) -> Callable[[Callable], Callable]:
    """ Registers the decorated creator under `name` and attaches `creator._check_submodule_params` (reference :133-160). """
    assert name not in submodule_creators, f'Error: "{name}" submodule creator entry already exists, can have duplicate submodule creator names.'
    if allowed_subm_params_keys is not None and required_subm_params_keys is not None:
        allowed_subm_params_keys = set(allowed_subm_params_keys) | set(required_subm_params_keys)

    def _decorator(creator: Callable[..., torch.nn.Module]):
        submodule_creators[name] = creator

        def _check_submodule_params(submodule_params: Dict[str, Any]):
            if allowed_subm_params_keys is not None:
                unknown = [n for n in submodule_params.keys() if n not in set(allowed_subm_params_keys)]
                if len(unknown) > 0:
                    raise ValueError(f'Error: "{unknown}" parameter(s) are not allowed for "{name}" creator throught `submodule_params` argument.{NL}'
                                     f' Allowed params are: "{allowed_subm_params_keys}"; Required params are "{required_subm_params_keys}"')
            if required_subm_params_keys is not None:
                missing = [n for n in set(required_subm_params_keys) if n not in submodule_params]
                if len(missing) > 0:
                    raise ValueError(f'Error: Missing "{missing}" parameter(s) in "{name}" creator\'s `submodule_params` params dict.{NL}'
                                     f'Required params are: "{required_subm_params_keys}"; Allowed params are: "{allowed_subm_params_keys}"')
        creator._check_submodule_params = _check_submodule_params
        return creator
    return _decorator


def _spatial_dims(input_shape) -> int:
    """ Shapes handed to creators are batch-less, channel first (reference nn_spec.py:104,173-175). """
    shape = input_shape if isinstance(input_shape[0], (int, np.integer)) else input_shape[0]
    return len(shape) - 1


@submodule_creator_dec(name='average_pooling')
def avg_pooling_creator(submodule_params: Dict[str, Any], input_shape) -> torch.nn.Module:
    """ reference :163-176 -> `nn.avg_pooling_nd` """
    return deepcv_nn.avg_pooling_nd(dims=_spatial_dims(input_shape), **submodule_params)


BASIC_SUBMODULE_CREATORS['avg_pooling'] = avg_pooling_creator  # the spelling conf/base/parameters.yml uses (:15,18,27...)


@submodule_creator_dec(name='reduce', allowed_subm_params_keys=set())
def reduction_subm_creator(submodule_params: Dict[str, Any], fn: str, keep_dim: bool = False) -> ForwardCallbackSubmodule:
    """ reference :179-186: standalone reduction of the (list of) tensor(s) coming from the previous submodule. """
    reduction_subm_creator._check_submodule_params(submodule_params)
    reduce = get_reduction_fn(fn)

    @deepcv_nn.forward_call_convention_dec(apply_parallel_forward=False, ignore_sub_refs=True)
    def _reduce_forward(tensors: List[torch.Tensor]):
        return reduce(tensors, keep_dim=keep_dim) if len(tensors) > 1 else tensors[0]
    return ForwardCallbackSubmodule(_reduce_forward)


def parse_slice(spec) -> Union[int, slice]:
    """ 'i' / 'start:stop[:step]' / int -> index or slice over a list of tensors (reference utils `parse_slice`). """
    if isinstance(spec, (int, np.integer)):
        return int(spec)
    parts = [int(v) if v.strip() else None for v in str(spec).split(':')]
    if len(parts) == 1:
        return parts[0]
    if len(parts) > 3:
        raise ValueError(f'Error: Invalid slice specification "{spec}"')
    return slice(*parts)


@submodule_creator_dec(name='select_tensor', required_subm_params_keys={'slice', })
def select_tensor_creator(submodule_params: Dict[str, Any], reduction: str = 'none') -> ForwardCallbackSubmodule:
    """ reference :188-200: keeps `tensors[slice]` of the parallel tensors coming from the previous submodule (then the optional reduction). """
    select_tensor_creator._check_submodule_params(submodule_params)
    parsed_slice = parse_slice(submodule_params['slice'])
    reduce = TENSOR_REDUCTION_FNS[reduction]

    @deepcv_nn.forward_call_convention_dec(apply_parallel_forward=False, ignore_sub_refs=True)
    def _select_tensor_forward(tensors: List[torch.Tensor]):
        picked = tensors[parsed_slice]
        return picked if deepcv_nn.is_torch_obj(picked) else reduce(picked)
    return ForwardCallbackSubmodule(_select_tensor_forward)


@submodule_creator_dec(name=NEW_BRANCH_FROM_TENSOR, allowed_subm_params_keys={FROM, FROM_NAS_INPUT_CHOICE})
def new_branch_creator(submodule_params: Dict[str, Any], reduction: str = 'concat') -> ForwardCallbackSubmodule:
    """ reference :203-224: ignores the previous output, continues from the (reduced) referenced tensor(s). """
    new_branch_creator._check_submodule_params(submodule_params)
    if FROM not in submodule_params and FROM_NAS_INPUT_CHOICE not in submodule_params:
        raise ValueError(f'Error: "{NEW_BRANCH_FROM_TENSOR}" submodules at least needs "{FROM}" or "{FROM_NAS_INPUT_CHOICE}" param in `submodule_params`')
    reduce = TENSOR_REDUCTION_FNS[reduction]

    @deepcv_nn.forward_call_convention_dec(apply_parallel_forward=True, ignore_prev_subm_intput=True, refs_tensor_count_similar=True)   # reference :210
    def _new_branch_forward(_prev_subm_out, referenced_submodules_out: List[torch.Tensor]):
        return reduce(list(referenced_submodules_out)) if len(referenced_submodules_out) > 1 else referenced_submodules_out[0]
    return ForwardCallbackSubmodule(_new_branch_forward)


def add_nn_layer_creator(layer_op_t: Type[torch.nn.Module], creator_name: str, submodule_creators: Dict[str, Callable] = BASIC_SUBMODULE_CREATORS) -> Callable:
    """ Registers a convolution / fully connected layer creator (reference :227-269): optional dropout, activation, normalisations,
    pre-activation order; `padding` defaults to `get_padding_from_kernel(kernel_size)`, `in_channels` / `in_features` to the input shape. """
    if not (deepcv_nn.is_conv(layer_op_t) or deepcv_nn.is_fully_connected(layer_op_t)):
        raise TypeError(f'Error: Wrong `layer_op_t` type, cant create a NN layer of type {layer_op_t} with `deepcv.meta.submodule_creators.add_nn_layer_creator` '
                        'submodule creator (`layer_op_t` should either be a convolution or a `torch.nn.Linear`).')

    @submodule_creator_dec(name=creator_name, submodule_creators=submodule_creators)
    def _nn_layer_creator(submodule_params: Dict[str, Any], input_shape, act_fn: Type[torch.nn.Module] = None, dropout_prob: float = None, preactivation: bool = False,
                          batch_norm=None, layer_norm=None, instance_norm=None, group_norm=None, layer_nrm_and_mean_batch_nrm=None) -> torch.nn.Module:
        submodule_params = dict(submodule_params)
        if not isinstance(input_shape[0], (int, np.integer)):   # parallel branches: ONE layer shared by all of them (reference :254), sized for the first
            if any(tuple(sh) != tuple(input_shape[0]) for sh in input_shape):
                raise ValueError(f'Error: parallel tensors fed to one "{creator_name}" layer must all have the same shape, got {list(input_shape)}')
            input_shape = input_shape[0]
        if deepcv_nn.is_fully_connected(layer_op_t):
            if 'in_features' not in submodule_params:
                submodule_params['in_features'] = int(np.prod(input_shape))
        else:
            if 'padding' not in submodule_params:
                submodule_params['padding'] = deepcv_nn.get_padding_from_kernel(submodule_params['kernel_size'], warn_on_uneven_kernel=False)
            if 'in_channels' not in submodule_params:
                submodule_params['in_channels'] = input_shape[0]
        if layer_op_t not in (torch.nn.Conv2d, torch.nn.Linear):
            raise NotImplementedError(f'deepcv_b200: "{creator_name}" ({layer_op_t.__name__}) is outside the sm_100a hot path; conv2d and linear / fully_connected are built')
        return deepcv_nn.layer(layer_op=layer_op_t(**submodule_params), act_fn=act_fn, dropout_prob=dropout_prob, preactivation=preactivation, input_shape=tuple(input_shape),
                               batch_norm=batch_norm, layer_norm=layer_norm, instance_norm=instance_norm, group_norm=group_norm, layer_nrm_and_mean_batch_nrm=layer_nrm_and_mean_batch_nrm)

    _nn_layer_creator.__doc__ = add_nn_layer_creator.__doc__
    return _nn_layer_creator


add_nn_layer_creator(layer_op_t=torch.nn.Conv1d, creator_name='conv1d')
add_nn_layer_creator(layer_op_t=torch.nn.Conv2d, creator_name='conv2d')
add_nn_layer_creator(layer_op_t=torch.nn.Conv3d, creator_name='conv3d')
add_nn_layer_creator(layer_op_t=torch.nn.ConvTranspose1d, creator_name='transosed_conv1d')
add_nn_layer_creator(layer_op_t=torch.nn.ConvTranspose2d, creator_name='transosed_conv2d')
add_nn_layer_creator(layer_op_t=torch.nn.ConvTranspose3d, creator_name='transosed_conv3d')
add_nn_layer_creator(layer_op_t=torch.nn.Linear, creator_name='linear')
add_nn_layer_creator(layer_op_t=torch.nn.Linear, creator_name='fully_connected')


def add_residual_dense_link_creator(is_residual: bool, creator_name: str, submodule_creators: Dict[str, Callable] = BASIC_SUBMODULE_CREATORS) -> Callable:
    """ Registers a residual (default reduction 'sum') or dense (default 'concat') link creator (reference :272-332): the previous
    output followed by every `_from` tensor — bilinearly rescaled to the previous output's spatial shape iff shapes differ and
    `allow_scaling` — is reduced. """
    @submodule_creator_dec(name=creator_name, submodule_creators=submodule_creators, allowed_subm_params_keys={FROM, FROM_NAS_INPUT_CHOICE})
    def _link_creator(submodule_params: Dict[str, Any], allow_scaling: bool = False, scaling_align_corners: bool = False, scaling_mode: str = None,
                      reduction: str = 'sum' if is_residual else 'concat', apply_in_parallel: bool = True, channel_dim: int = 1) -> ForwardCallbackSubmodule:
        _link_creator._check_submodule_params(submodule_params)
        if FROM not in submodule_params and FROM_NAS_INPUT_CHOICE not in submodule_params:
            raise ValueError(f'Error: Missing "{FROM}" or "{FROM_NAS_INPUT_CHOICE}" parameter in '
                             f'{creator_name} link YAML specification; You should at least provide a tensor reference.')
        reduce = TENSOR_REDUCTION_FNS[reduction]

        @deepcv_nn.forward_call_convention_dec(apply_parallel_forward=apply_in_parallel, in_tensors_count_similar_to_refs=apply_in_parallel)   # reference :300
        def _forward_callback(x, referenced_submodules_out: List[torch.Tensor]):
            if isinstance(x, ops.PendingAffine) or (isinstance(x, ops.PendingNorm) and reduction != 'sum'):
                x = ops.materialize(x)
            out = [x] if deepcv_nn.is_torch_obj(x) else [ops.materialize(t) for t in x]
            linked = [y for refs in referenced_submodules_out for y in ([refs] if isinstance(refs, torch.Tensor) else list(refs))]
            if reduction == 'concat' and allow_scaling and channel_dim == 1 and not scaling_align_corners and scaling_mode in (None, 'bilinear'):
                fused = ops.link_concat_rescaled(out + linked)   # same-size and exactly-2x references: rescale + concat in one launch
                if fused is not None:
                    return fused
            for y in linked:
                if out[0].shape[channel_dim + 1:] != y.shape[channel_dim + 1:]:
                    if not allow_scaling:
                        raise RuntimeError(f"Error: Couldn't forward throught {creator_name} link: features from link doesn't have "
                                           f"the same shape as previous module's output shape, can't concatenate or add them. (did you forgot to allow residual/dense "
                                           f"features to be scaled using `allow_scaling: true` parameter?). `residual_shape='{y.shape}' != prev_features_shape='{out[0].shape}'`")
                    y = deepcv_nn.interpolate(y, out[0].shape[channel_dim + 1:], scaling_mode=scaling_mode, align_corners=scaling_align_corners)
                out.append(y)
            return reduce(out)
        subm = ForwardCallbackSubmodule(_forward_callback)
        # a block's raw output with its normalisation pending (`ops.PendingNorm`) may be this link's first operand: a 'sum' adds the referenced tensor inside
        # the block's apply pass (ops.link_reduce); anything else materialises it first
        subm.accepts_pending_affine = subm.accepts_pending_with_references = True
        return subm

    _link_creator.__doc__ = add_residual_dense_link_creator.__doc__
    return _link_creator


add_residual_dense_link_creator(is_residual=True, creator_name='residual_link')
add_residual_dense_link_creator(is_residual=False, creator_name='dense_link')


def _outside_hot_path(name: str) -> Callable:
    def _creator(submodule_params: Dict[str, Any] = None):
        raise NotImplementedError(f'deepcv_b200: the "{name}" submodule is outside the DeepcvModule conv/BN/augment hot path and is not built for sm_100a')
    return _creator


for _name in ('concat_coords', 'concat_hilbert_coords', 'multiresolution_fusion', 'parallel_conv', 'hrnet_input_stem', 'hrnet_repr_head_v1', 'hrnet_repr_head_vZ',
              'hrnet_repr_head_v2p'):
    BASIC_SUBMODULE_CREATORS[_name] = _outside_hot_path(_name)
