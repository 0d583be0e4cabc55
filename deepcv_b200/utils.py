""" The few helpers of `src/deepcv/utils.py` the hot path's boundary needs: `get_by_identifier` (:365-379, used by
`nn_spec.py:164` and `preprocess.py:162` to turn dotted strings into callables) and `set_seeds` (:65-84). """
import importlib
import random
import re
from typing import Optional

import numpy as np
import torch

__all__ = ['get_by_identifier', 'set_seeds', 'NL']
NL = '\n'


def get_by_identifier(identifier: str):
    regex = r'[\w\.]*\w'
    if not isinstance(identifier, str) or not re.fullmatch(regex, identifier):
        raise ValueError(f'Error: bad identifier given in `deepcv.utils.get_by_identifier` function (identifier="{identifier}" must match "{regex}" regex)')
    *module_parts, name = identifier.split('.')
    if module_parts:
        module = importlib.import_module('.'.join(module_parts))
        return getattr(module, name)
    if name in globals():
        return globals()[name]
    raise RuntimeError(f'Error: can\'t find ``{identifier}`` identifier (you may have to specify its module)')


def set_seeds(seed: int = 345349, set_different_seeds: bool = True) -> None:
    """ torch / CUDA / numpy / python RNG seeds from one seed: increments of `seed` when `set_different_seeds` (reference :65-84). """
    torch_seed, cuda_seed, np_seed, python_seed = (range(seed, seed + 4) if set_different_seeds else [seed] * 4)
    torch.manual_seed(torch_seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(cuda_seed)
    np.random.seed(np_seed % (2 ** 32))
    random.seed(python_seed)
