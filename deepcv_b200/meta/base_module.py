""" `DeepcvModule` — host-side mirror of `src/deepcv/meta/base_module.py:39-264`.

Constructor signature, `HP_DEFAULTS`, `SUBM_CREATOR_SPECIAL_ARGS`, the YAML-driven construction, the named-tensor-reference
forward and the Xavier initialisation are the reference's; the submodules execute the sm_100a kernels (`deepcv_b200.ops`).
Differences that follow from "no CPU fallback on the named path":
  * `forward` accepts CUDA tensors (float32 = parity mode, bfloat16 = throughput mode; N x C x H x W, any memory format — NHWC
    is used internally) and `meta` tensors (shape inference); CPU tensors raise.
  * `weight_norm` / `spectral_norm` hyper-parameters must be None (the reference applies them to a container without a `weight`
    and raises).
"""
import copy
from collections import OrderedDict
from functools import partial
from typing import Any, Callable, Dict, List, Optional, Sequence, Type, Union

import numpy as np
import torch

from .. import ops
from . import nn as deepcv_nn
from . import nn_spec
from .hyperparams import HYPERPARAMS_T, to_hyperparameters

__all__ = ['DeepcvModule', 'DeepcvModuleDescriptor']
NL = '\n'


class DeepcvModule(torch.nn.Module):
    """ A `torch.nn.Module` whose architecture is declared by the `architecture` hyper-parameter (a YAML list, see
    conf/base/parameters.yml) and built through the submodule-creator registry (`deepcv_b200.meta.submodule_creators`). """

    HP_DEFAULTS = {'architecture': ..., 'act_fn': ..., 'weight_norm': None, 'spectral_norm': None}
    SUBM_CREATOR_SPECIAL_ARGS = {'submodule_params', 'prev_shapes', 'input_shape', 'input_shapes'}

    def __init__(self, input_shape, hp: HYPERPARAMS_T, additional_submodule_creators: Optional[Dict[str, Union[Callable, Type[torch.nn.Module]]]] = None,
                 extend_basic_submodule_creators_dict: bool = True, additional_init_logic: Callable[[torch.nn.Module, Any], None] = None):
        super().__init__()
        self._input_shape = tuple(input_shape)
        self._single_input_shape = isinstance(input_shape[0], (int, np.integer))
        if not self._single_input_shape:
            raise NotImplementedError('deepcv_b200: multiple input tensors (parallel branches from the input) are outside the hot path')
        self._spatial_dims = len(input_shape[1:])
        self._uses_nni_nas_mutables = False
        self._uses_forward_callback_submodules = False
        self._additional_init_logic = additional_init_logic

        assert self.HP_DEFAULTS != ..., f'Error: Module classes which inherits from "DeepcvModule" ({type(self).__name__}) must define "HP_DEFAULTS" class attribute dict.'
        self._hp, _missing = to_hyperparameters(hp, defaults=self.HP_DEFAULTS, raise_if_missing=True)

        nn_spec.define_nn_architecture(self, self._hp['architecture'], submodule_creators=additional_submodule_creators,
                                       extend_basic_submodule_creators_dict=extend_basic_submodule_creators_dict)
        self._initialize_parameters(self._hp['act_fn'], additional_init_logic=self._additional_init_logic)

        if self._hp['weight_norm'] is not None or self._hp['spectral_norm'] is not None:
            raise NotImplementedError('deepcv_b200: `weight_norm` / `spectral_norm` are not built for sm_100a (set them to null; the reference raises on them too)')

    # ---- forward (reference :113-155) -------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            raise ValueError(f'Error: No enought input tensors in "{type(self).__name__}", was expecting a single tensor of shape (N, {self._input_shape})')
        if x.device.type not in ('cuda', 'meta'):
            raise RuntimeError(f'deepcv_b200: DeepcvModule.forward got a tensor on "{x.device}"; this path runs on CUDA (sm_100a) only and has no CPU fallback')
        if x.dtype == torch.uint8:
            raise TypeError('deepcv_b200: got a uint8 batch; apply the preprocess recipe (`deepcv_b200.meta.data.preprocess.FusedPreprocess`) first')
        if x.device.type == 'cuda' and x.dim() == 4:
            x = ops.as_nhwc(x)

        # A few-channel convolution block may hand its RAW output on with the normalisation still pending (`ops.PendingAffine`) when the next
        # submodule applies it while loading (another such block, an average pooling) and nobody else refers to that output.
        names = list(self._child_modules._modules.keys()) if isinstance(self._child_modules, torch.nn.Sequential) else list(self._child_modules.keys())
        children = [self._child_modules._modules[n] for n in names] if isinstance(self._child_modules, torch.nn.Sequential) else [self._child_modules[n] for n in names]
        sequential = self.is_sequential_nn()
        referenced_output_features: Dict[str, torch.Tensor] = {}
        remaining_submodule_references = {} if sequential else copy.copy(self._submodule_references)
        for i, (name, subm) in enumerate(zip(names, children)):
            if isinstance(x, (ops.PendingAffine, ops.PendingNorm)) and not getattr(subm, 'accepts_pending_affine', False):
                x = ops.materialize(x)
            if isinstance(x, ops.PendingFlatten) and not getattr(subm, 'accepts_pending_flatten', False):
                x = ops.materialize(x)
            n_referrers = sum(name in r for r in remaining_submodule_references.values())
            kwargs = {}
            if getattr(subm, 'can_defer_affine', False):
                # ... or a residual link that sums this block's output with a referenced tensor (`accepts_pending_with_references`)
                kwargs['defer_affine'] = not isinstance(x, (list, tuple)) and i + 1 < len(children) and n_referrers == 0 and getattr(children[i + 1], 'accepts_pending_affine', False) \
                    and (not getattr(children[i + 1], 'referenced_submodules', None) or getattr(children[i + 1], 'accepts_pending_with_references', False))
            if getattr(subm, 'can_defer_flatten', False):   # Flatten in front of a small fully connected head: fused into the head's kernels
                kwargs['defer_flatten'] = not isinstance(x, (list, tuple)) and i + 1 < len(children) and n_referrers == 0 \
                    and bool(getattr(children[i + 1], 'accepts_pending_flatten', False)) and not getattr(children[i + 1], 'referenced_submodules', None)
            refs = None if sequential else getattr(subm, 'referenced_submodules', None)
            if refs is not None and len(refs) > 0:
                current_subm_references = OrderedDict([(ref, referenced_output_features[ref]) for ref in refs])
                # Release stored features once their last referrer has consumed them
                del remaining_submodule_references[name]
                for referenced_submodule in refs:
                    if not any(referenced_submodule in r for r in remaining_submodule_references.values()):
                        referenced_output_features.pop(referenced_submodule, None)
                x = subm(x, referenced_submodules_out=current_subm_references, **kwargs)
            else:
                x = subm(x, **kwargs)
            # Keep this output if a later submodule names it in `_from`. The kept tensor and the one flowing on are two autograd
            # aliases whose gradients are summed by a library kernel (ops.fork).
            n_referrers = sum(name in r for r in remaining_submodule_references.values())
            if n_referrers > 0:
                if isinstance(x, (list, tuple)):   # parallel branches: every tensor of the list is kept and flows on
                    pairs = [ops.fork(t) if t.device.type == 'cuda' else (t, t) for t in (ops.materialize(t) for t in x)]
                    x, referenced_output_features[name] = [a for a, _ in pairs], [b for _, b in pairs]
                else:
                    x = ops.materialize(x)
                    x, referenced_output_features[name] = ops.fork(x) if x.device.type == 'cuda' else (x, x)
        return [ops.materialize(t) for t in x] if isinstance(x, (list, tuple)) else ops.materialize(x)

    def __str__(self) -> str:
        return str(self.describe())

    def uses_nni_nas_mutables(self, recursive: bool = False) -> bool:
        if not recursive:
            return self._uses_nni_nas_mutables
        return self._uses_nni_nas_mutables or any(subm.uses_nni_nas_mutables(recursive=True) for subm in self._submodules.values() if isinstance(subm, DeepcvModule))

    def uses_forward_callback_submodules(self, recursive: bool = False) -> bool:
        if not recursive:
            return self._uses_forward_callback_submodules
        return self._uses_forward_callback_submodules or any(subm.uses_forward_callback_submodules(recursive=True) for subm in self._submodules.values() if isinstance(subm, DeepcvModule))

    def is_sequential_nn(self, recursive: bool = False) -> bool:
        """ True iff no submodule uses tensor references: `_child_modules` is then a `torch.nn.Sequential` (reference :179-182). """
        return not self.uses_nni_nas_mutables(recursive=recursive) and not self.uses_forward_callback_submodules(recursive=recursive)

    def describe(self):
        return DeepcvModuleDescriptor(self)

    # ---- initialisation (reference :230-264) --------------------------------------------------------------------------------------
    def _initialize_parameters(self, act_fn: Type[torch.nn.Module] = None, additional_init_logic: Callable[[torch.nn.Module, Any], None] = None):
        """ Xavier initialisation with the gain of the (global) activation function: convolutions `xavier_normal_`, fully connected
        `xavier_uniform_`, biases 0, BatchNorm gamma ~ U(0,1) and beta 0. Applied through `self.apply`, so an outer module re-initialises
        its nested ones with the outer gain, like the reference. Other parameterised modules (e.g. GroupNorm) keep their own defaults unless
        `additional_init_logic` handles them (the reference raises `TypeError` for them when no callback is given; the shipped
        `basic_backbone` spec, which uses `group_norm`, would therefore not construct — SURVEY.md section 8.c.2). """
        xavier_gain = torch.nn.init.calculate_gain(deepcv_nn.get_gain_name(act_fn)) if act_fn else None

        def _raise_if_no_act_fn(sub_module_name: str):
            if xavier_gain is None:
                raise ValueError(f'Error: Must specify `act_fn` argument in `DeepcvModule._initialize_parameters` function in order to initialize '
                                 f'{sub_module_name} sub-module(s) with Xavier Init. (See `deepcv.meta.nn.get_gain_name` for supported activation functions)')

        def _xavier_init(module: torch.nn.Module, additional_init):
            if deepcv_nn.is_conv(module):
                _raise_if_no_act_fn('convolution')
                torch.nn.init.xavier_normal_(module.weight.data, gain=xavier_gain)
                if module.bias is not None:
                    module.bias.data.fill_(0.)
            elif deepcv_nn.is_fully_connected(module):
                _raise_if_no_act_fn('fully connected')
                torch.nn.init.xavier_uniform_(module.weight.data, gain=xavier_gain)
                if module.bias is not None:
                    module.bias.data.fill_(0.)
            elif type(module).__module__ == torch.nn.BatchNorm2d.__module__:
                if module.weight is not None:
                    torch.nn.init.uniform_(module.weight.data)
                    module.bias.data.fill_(0.)
            elif len(list(module.parameters(recurse=False))) > 0 and additional_init is not None:
                additional_init(module, xavier_gain=xavier_gain)
        self.apply(partial(_xavier_init, additional_init=additional_init_logic))


class DeepcvModuleDescriptor:
    """ Human-readable description of a `DeepcvModule` (reference :362-415): capacity, submodule names, feature shapes. """

    def __init__(self, module: DeepcvModule):
        self.module = module
        self.capacity = deepcv_nn.get_model_capacity(module)
        self.human_readable_capacity = f'{self.capacity:,}'
        self.model_class, self.model_class_name = type(module), type(module).__name__
        if hasattr(module, '_features_shapes'):
            self.submodules_features_shapes = module._features_shapes
            self.submodules_features_dims = [len(s) for s in module._features_shapes]
        if hasattr(module, '_submodules_capacities'):
            self.submodules_capacities = module._submodules_capacities
            self.human_readable_capacities = [f'{c:,}' for c in module._submodules_capacities]
        if hasattr(module, '_architecture_spec'):
            self.architecture = module._architecture_spec
            self.submodules = {n: str(m) for n, m in module._submodules.items()}
        if hasattr(type(module), '__doc__'):
            self.model_class_docstring = type(module).__doc__

    def __str__(self) -> str:
        if hasattr(self, 'architecture'):
            features = self.submodules_features_shapes[1:]
            capas = self.human_readable_capacities
            names = list(self.submodules.keys())
            desc_str = f'{NL}\t- '.join(f'{n}({capa}) output_features_shape={s}' for n, capa, s in zip(names, capas, features))
        else:
            desc_str = '(No submodule architecture informations to describe: probably not a `DeepcvModule`)'
        return f'{self.model_class_name} (capacity={self.human_readable_capacity}):{NL}\t- {desc_str}'
