""" YAML configuration loading for `parameters.yml`-style files.

The reference loads `conf/base/parameters.yml` through Kedro's ConfigLoader -> anyconfig -> ruamel.yaml with the
"unsafe" loader so that `!py!torch.nn.ReLU` tags become Python objects (reference: `src/deepcv/run.py:40-51`,
`src/deepcv/utils.py:55-62`, tag directives at `conf/base/parameters.yml:1-3`). Neither kedro, anyconfig nor ruamel is
needed here: this module gives PyYAML the same behaviour for the two tag prefixes the reference uses and for YAML 1.2
floats written without a dot (`1e-05`, `1e-3`), which PyYAML's YAML 1.1 resolver would otherwise read as strings.
"""
import importlib
import re
from pathlib import Path
from typing import Any, Dict, Union

import yaml

__all__ = ['load_parameters', 'loads_parameters', 'UnresolvedPythonName', 'PYTHON_NAME_SHIMS', 'find_model_spec', 'benchmark_model_spec']

_PY_NAME_PREFIX = 'tag:yaml.org,2002:python/name:'
_PY_OBJECT_PREFIX = 'tag:yaml.org,2002:python/object:'

""" Dotted names which live in packages absent from this image are mapped to in-tree equivalents (same call signature). """
PYTHON_NAME_SHIMS: Dict[str, str] = {
    'ignite.contrib.handlers.PiecewiseLinear': 'deepcv_b200.meta.ignite_training.PiecewiseLinear',
    'ignite.handlers.PiecewiseLinear': 'deepcv_b200.meta.ignite_training.PiecewiseLinear',
}


class UnresolvedPythonName:
    """ Placeholder for a `!py!` tag whose module cannot be imported here (e.g. a `deepcv.*` or `nni.*` name).
    Loading the file must not fail because an unrelated section names a missing package; using the placeholder does. """

    def __init__(self, dotted_name: str, error: Exception):
        self.dotted_name, self.error = dotted_name, error

    def __call__(self, *args, **kwargs):
        raise ImportError(f'Error: "{self.dotted_name}" (from a `!py!` YAML tag) could not be resolved: {self.error}')

    def __repr__(self):
        return f'UnresolvedPythonName({self.dotted_name!r})'


def resolve_python_name(dotted_name: str) -> Any:
    dotted_name = PYTHON_NAME_SHIMS.get(dotted_name, dotted_name)
    module_name, _, attr = dotted_name.rpartition('.')
    try:
        if not module_name:
            import builtins
            return getattr(builtins, attr)
        # Longest importable module prefix, then attribute walk (handles `torch.nn.ReLU` and nested classes alike)
        parts = dotted_name.split('.')
        for split in range(len(parts) - 1, 0, -1):
            try:
                obj = importlib.import_module('.'.join(parts[:split]))
            except ImportError:
                continue
            for name in parts[split:]:
                obj = getattr(obj, name)
            return obj
        raise ImportError(f'no importable prefix in "{dotted_name}"')
    except (ImportError, AttributeError) as e:
        return UnresolvedPythonName(dotted_name, e)


class ParametersLoader(yaml.SafeLoader):
    """ SafeLoader + python/name tags + YAML 1.2 float forms. Nothing else from PyYAML's unsafe constructors is enabled. """


def _construct_python_name(loader: yaml.Loader, suffix: str, node: yaml.Node) -> Any:
    value = loader.construct_scalar(node) if isinstance(node, yaml.ScalarNode) else None
    if value:
        raise yaml.constructor.ConstructorError(None, None, f'expected the empty value for a python/name tag, found {value!r}', node.start_mark)
    return resolve_python_name(suffix)


ParametersLoader.add_multi_constructor(_PY_NAME_PREFIX, _construct_python_name)
ParametersLoader.add_multi_constructor(_PY_OBJECT_PREFIX, _construct_python_name)
ParametersLoader.add_implicit_resolver(
    'tag:yaml.org,2002:float',
    re.compile(r'''^(?:[-+]?[0-9][0-9_]*\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                   |[-+]?[0-9][0-9_]*[eE][-+]?[0-9]+
                   |[-+]?\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                   |[-+]?\.(?:inf|Inf|INF)
                   |\.(?:nan|NaN|NAN))$''', re.X),
    list('-+0123456789.'))


def _hashable_keys_fix(loader, node, deep=False):
    # `!py!torchvision.transforms.Normalize "": {...}` uses a Python type as a mapping key (parameters.yml:189,201,210)
    loader.flatten_mapping(node)
    mapping = {}
    for key_node, value_node in node.value:
        key = loader.construct_object(key_node, deep=True)
        mapping[key] = loader.construct_object(value_node, deep=deep)
    return mapping


ParametersLoader.construct_mapping = _hashable_keys_fix


def loads_parameters(text: str) -> Dict[str, Any]:
    return yaml.load(text, Loader=ParametersLoader)


def load_parameters(path: Union[str, Path]) -> Dict[str, Any]:
    with open(path, 'r') as f:
        return loads_parameters(f.read())


def find_model_spec(parameters: Dict[str, Any], model_name: str) -> Dict[str, Any]:
    """ `models:` is a YAML *list* of single-key mappings (`parameters.yml:6-98`); returns the hp dict of the named one. """
    for entry in parameters['models']:
        if model_name in entry:
            return entry[model_name]
    raise KeyError(f'Error: no model named "{model_name}" under `models:` (found: {[next(iter(e)) for e in parameters["models"]]})')


def benchmark_model_spec(path: Union[str, Path], model_name: str, out_features: int = None) -> Dict[str, Any]:
    """ The hp dict of `model_name` from a `parameters.yml`-style file, as the benchmark configs use it (note iii):
    a private copy with `spectral_norm: null` — `conf/base/parameters.yml:83` sets it and `base_module.py:109-111` then applies
    `torch.nn.utils.spectral_norm` to a container without a `weight`, which raises — and, optionally, the head's `out_features`
    (the reference deduces it from the dataset's classes, `classification/image.py:44-50`). """
    import copy
    hp = dict(find_model_spec(load_parameters(path), model_name))
    hp['architecture'] = copy.deepcopy(hp['architecture'])
    hp['spectral_norm'] = None
    if out_features is not None:
        hp['architecture'][-1]['fully_connected']['out_features'] = int(out_features)
    return hp
