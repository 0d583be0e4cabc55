""" YAML architecture-specification parser — host-side mirror of `src/deepcv/meta/nn_spec.py`.

Grammar, tokens, parameter merging and the creator calling convention are the reference's (`yaml_tokens` :35-50,
`define_nn_architecture` :55-104, `_parse_torch_module_from_submodule_spec` :107-191, `_subm_name_and_params_from_spec`
:194-215, `_setup_forward_callback_submodule` :218-243), with its defect ledger applied (tokens compared
by value, both spellings of the nested-module token, `isinstance(x, dict)`, references looked up with `.get`). NNI NAS mutables
(`_nas_layer_choice`, `_from_nas_input_choice`) belong to the AutoML subsystem, not to this path: naming them raises.

Shape inference after each submodule uses a `meta`-device forward (`nn.get_out_features_shape`), so building a model launches no kernel.
"""
import copy
import enum
import inspect
from collections import OrderedDict
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Type, Union

import numpy as np
import torch

from .. import utils
from . import nn as deepcv_nn
from .nn import get_model_capacity, get_out_features_shape

__all__ = ['DEFAULT_LAYER_CHOICE_REDUCTION', 'yaml_tokens', 'define_nn_architecture', 'TYPE_SUBSTITUTIONS']

DEFAULT_LAYER_CHOICE_REDUCTION = r'mean'


class yaml_tokens(str, enum.Enum):
    """ Special tokens of the YAML architecture specification (reference :35-50). Members compare equal to their string value. """
    FROM = r'_from'
    SUBMODULE_NAME = r'_name'
    NAS_LAYER_CHOICE = '_nas_layer_choice'
    NESTED_DEEPCV_MODULE = r'_nested_deepcv_module'
    FROM_NAS_INPUT_CHOICE = r'_from_nas_input_choice'
    NAS_LAYER_REDUCTION_FN = r'_reduction'
    NEW_BRANCH_FROM_TENSOR = '_new_branch_from_tensor'
    FROM_NAS_INPUT_N_CHOSEN = r'_n_chosen'
    NAS_MUTABLE_RETURN_MASK = r'_return_mask'
    NAS_LAYER_CHOICE_CANDIDATES = r'_candidates'

    def __str__(self):
        return self.value


# `conf/base/parameters.yml:85` spells the nested-module token without the second underscore
_NESTED_TOKENS = (yaml_tokens.NESTED_DEEPCV_MODULE.value, '_nested_deepcvmodule')

""" Stock `torch.nn` types a spec may name directly (`!py!torch.nn.Flatten`, parameters.yml:87) whose work must stay on the
library's kernels: the parser instantiates the mapped subclass instead (same constructor, same state_dict). """
TYPE_SUBSTITUTIONS: Dict[Type[torch.nn.Module], Type[torch.nn.Module]] = {torch.nn.Flatten: deepcv_nn.Flatten, torch.nn.AvgPool2d: deepcv_nn.AvgPool2d}


def define_nn_architecture(deepcv_module, architecture_spec: Iterable, submodule_creators: Dict[str, Callable] = None, extend_basic_submodule_creators_dict: bool = True):
    """ Parses `architecture_spec` and creates `deepcv_module`'s submodules (reference :55-104). Defines `_features_shapes`,
    `_submodules_capacities`, `_submodules`, `_architecture_spec`, `_submodule_references` and `_child_modules`. """
    from .submodule_creators import BASIC_SUBMODULE_CREATORS, ForwardCallbackSubmodule
    deepcv_module._features_shapes = [deepcv_module._input_shape]
    deepcv_module._architecture_spec = architecture_spec
    deepcv_module._submodules_capacities = list()
    deepcv_module._submodules = OrderedDict()
    # referrer sub-module name -> names of the sub-modules whose output it consumes (`_from`)
    deepcv_module._submodule_references: Dict[str, List[str]] = dict()

    subm_creators = {**(BASIC_SUBMODULE_CREATORS if extend_basic_submodule_creators_dict else dict()), **(submodule_creators if submodule_creators is not None else dict())}
    deepcv_module._subm_creators = subm_creators

    for i, submodule_spec in enumerate(architecture_spec):
        subm_name, subm = _parse_torch_module_from_submodule_spec(deepcv_module, submodule_spec, i, subm_creators)
        deepcv_module._submodules[subm_name] = subm
        if isinstance(subm, ForwardCallbackSubmodule) and getattr(subm, 'referenced_submodules', None) is not None:
            deepcv_module._submodule_references[subm_name] = subm.referenced_submodules
        deepcv_module._submodules_capacities.append(get_model_capacity(subm))
        if deepcv_module.is_sequential_nn():
            deepcv_module._child_modules = torch.nn.Sequential(deepcv_module._submodules)
        else:
            deepcv_module._child_modules = torch.nn.ModuleDict(deepcv_module._submodules)
        missing_refs = [ref for ref in deepcv_module._submodule_references.get(subm_name, []) if ref not in deepcv_module._submodules.keys()]
        if len(missing_refs) > 0:
            raise ValueError(f'Error: Invalid sub-module reference(s), cant find following sub-module name(s)/label(s): "{missing_refs}".'
                             ' Output tensor references must refer to a previously defined sub-module name.')
        # Output shape of the new submodule: dummy (meta-device) forward of the submodule on the previous shape(s)
        deepcv_module._features_shapes.append(_infer_out_shape(deepcv_module, subm_name, subm))


def _infer_out_shape(deepcv_module, subm_name: str, subm: torch.nn.Module):
    """ The reference re-runs the *whole* module on zeros after each addition (O(L^2), :103-104); one submodule on a `meta` tensor of
    the previous shape gives the same answer. Link submodules get the recorded shapes of the tensors they reference. """
    def _dummy(shape):   # a shape, or a list of shapes when the submodule produced several tensors (parallel branches)
        return deepcv_nn.meta_like((1, *shape)) if isinstance(shape[0], (int, np.integer)) else [deepcv_nn.meta_like((1, *s)) for s in shape]
    x = _dummy(deepcv_module._features_shapes[-1])
    was_training = subm.training
    with torch.no_grad():
        refs = deepcv_module._submodule_references.get(subm_name)
        if refs:
            names = list(deepcv_module._submodules.keys())
            ref_out = OrderedDict((r, _dummy(deepcv_module._features_shapes[names.index(r) + 1])) for r in refs)
            out = subm(x, referenced_submodules_out=ref_out)
        else:
            out = subm(x)
    subm.train(was_training)
    return tuple(out.shape[1:]) if isinstance(out, torch.Tensor) else [tuple(t.shape[1:]) for t in out]


def _parse_torch_module_from_submodule_spec(deepcv_module, submodule_spec, submodule_pos: Union[int, str], subm_creators: Dict[str, Callable],
                                            default_submodule_prefix: str = '_submodule_', allow_mutable_layer_choices: bool = True) -> Tuple[str, torch.nn.Module]:
    """ One submodule from its spec (reference :107-191). """
    from .submodule_creators import ForwardCallbackSubmodule
    subm_name = default_submodule_prefix + str(submodule_pos)
    subm_name, params, subm_type = _subm_name_and_params_from_spec(submodule_spec, default_subm_name=subm_name, existing_subm_names=deepcv_module._submodules.keys())

    # Global (hyper)parameters of `hp` are defaults for every submodule; local spec entries override them
    params_with_globals = {n: copy.deepcopy(v) for n, v in deepcv_module._hp.items() if n not in params}
    params_with_globals.update(params)

    if isinstance(subm_type, str) and subm_type in _NESTED_TOKENS:
        module = type(deepcv_module)(input_shape=deepcv_module._features_shapes[-1], hp=params_with_globals, additional_submodule_creators=subm_creators,
                                     extend_basic_submodule_creators_dict=False, additional_init_logic=deepcv_module._additional_init_logic)
    elif isinstance(subm_type, str) and subm_type == yaml_tokens.NAS_LAYER_CHOICE:
        raise NotImplementedError(f'deepcv_b200: "{yaml_tokens.NAS_LAYER_CHOICE}" (NNI NAS LayerChoice) belongs to the AutoML subsystem, which is outside the DeepcvModule hot path')
    else:
        if isinstance(subm_type, str):
            fn_or_type = subm_creators.get(subm_type)
            if not fn_or_type:
                try:
                    fn_or_type = utils.get_by_identifier(subm_type)
                except Exception as e:
                    raise RuntimeError(f'Error: Could not locate module/function named "{subm_type}" given module creators: "{subm_creators.keys()}"') from e
        else:
            fn_or_type = subm_type
        fn_or_type = TYPE_SUBSTITUTIONS.get(fn_or_type, fn_or_type) if isinstance(fn_or_type, type) else fn_or_type
        if not callable(fn_or_type):
            raise RuntimeError(f'Error: Invalid sub-module creator function or type: "{fn_or_type}"')

        submodule_signature_params = inspect.signature(fn_or_type).parameters
        params_with_globals['prev_shapes'] = deepcv_module._features_shapes
        params_with_globals['input_shape'] = deepcv_module._features_shapes[-1]
        params_with_globals['input_shapes'] = deepcv_module._features_shapes[-1]
        provided_params = {n: v for n, v in params_with_globals.items() if n in submodule_signature_params}
        if 'submodule_params' in submodule_signature_params:
            provided_params['submodule_params'] = {n: v for n, v in params.items() if n not in provided_params}
        module = fn_or_type(**provided_params)

        if isinstance(module, ForwardCallbackSubmodule):
            _setup_forward_callback_submodule(deepcv_module, subm_name, submodule_params=params, forward_callback_module=module)
        elif not isinstance(module, torch.nn.Module):
            raise RuntimeError('Error: Invalid sub-module creator function or type: '
                               'Must either be a `torch.nn.Module` (Type or string identifier of a Type) or a submodule creator which returns a `torch.nn.Module`.')
    return subm_name, module


def _subm_name_and_params_from_spec(submodule_spec, default_subm_name: str, existing_subm_names: Sequence[str]) -> Tuple[str, Dict, Union[Type[torch.nn.Module], str, Callable]]:
    """ (name, params, type) for every spelling of a submodule spec (reference :194-215): `type`, `{type: params}`,
    `{type: [name, params]}`, `{type: "name"}`, `{type: {_name: name, ...}}`. """
    subm_type, params = list(submodule_spec.items())[0] if isinstance(submodule_spec, dict) else (submodule_spec, {})
    subm_name = default_subm_name
    if isinstance(params, (list, tuple)):
        subm_name, params = params[0], params[1]
    elif isinstance(params, str):
        subm_name, params = params, dict()
    elif isinstance(params, dict) and yaml_tokens.SUBMODULE_NAME.value in params:
        params = dict(params)
        subm_name = params.pop(yaml_tokens.SUBMODULE_NAME.value)
    if params is None:
        params = dict()
    if subm_name in existing_subm_names or subm_name == r'' or not isinstance(subm_name, str):
        raise ValueError(f'Error: Invalid or duplicate sub-module name/label: "{subm_name}"')
    if not isinstance(params, dict):
        raise RuntimeError(f'Error: Architecture sub-module spec. must either be a parameters Dict, or a submodule name along with parameters Dict, but got: "{params}".')
    return subm_name, dict(params), subm_type


def _setup_forward_callback_submodule(deepcv_module, subm_name: str, submodule_params: Dict[str, Any], forward_callback_module) -> None:
    """ Records the tensor references of a `ForwardCallbackSubmodule` (reference :218-243). """
    deepcv_module._uses_forward_callback_submodules = True
    FROM, CHOICE = yaml_tokens.FROM.value, yaml_tokens.FROM_NAS_INPUT_CHOICE.value
    if CHOICE in submodule_params:
        raise NotImplementedError(f'deepcv_b200: "{CHOICE}" (NNI NAS InputChoice) belongs to the AutoML subsystem, which is outside the DeepcvModule hot path')
    if yaml_tokens.NAS_MUTABLE_RETURN_MASK.value in submodule_params or yaml_tokens.FROM_NAS_INPUT_N_CHOSEN.value in submodule_params:
        raise ValueError(f'Error: Cannot specify "{yaml_tokens.NAS_MUTABLE_RETURN_MASK}" nor "{yaml_tokens.FROM_NAS_INPUT_N_CHOSEN}" without using "{CHOICE}".')
    if FROM in submodule_params:
        tensor_references = submodule_params[FROM]
        forward_callback_module.referenced_submodules = [tensor_references] if isinstance(tensor_references, str) else list(tensor_references)
