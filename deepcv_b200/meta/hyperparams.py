""" Hyper-parameter mapping with required/default handling.

Mirrors the two names the hot path's boundary uses from the reference: `Hyperparameters`
(`src/deepcv/meta/data/training_metadata.py:61-118`) and `to_hyperparameters` (`src/deepcv/meta/hyperparams.py:229-248`).
A value of `...` (Ellipsis) in `defaults` marks a required entry. The reference's meta-learning classes that live next to
these (embeddings, scale predictors) are not part of the path and are not built.
"""
import collections.abc
import logging
import types
from typing import Any, Dict, List, Mapping, Tuple, Union

__all__ = ['Hyperparameters', 'to_hyperparameters', 'HYPERPARAMS_T']


class Hyperparameters(collections.abc.Mapping):
    """ Read-only mapping of hyper-parameters (hashable when its values are). """

    def __init__(self, **kwargs):
        self._store = dict(**kwargs)
        self._hash = None

    def __iter__(self):
        return iter(self._store)

    def __len__(self):
        return len(self._store)

    def __getitem__(self, key):
        return self._store[key]

    def __hash__(self):
        if self._hash is None:
            h = 0
            for pair in self.items():
                h ^= hash(repr(pair))
            self._hash = h
        return self._hash

    def __eq__(self, other):
        if isinstance(other, Hyperparameters):
            return self._store == other._store
        if isinstance(other, collections.abc.Mapping):
            return self._store == dict(other)
        return NotImplemented

    def __repr__(self):
        return f'Hyperparameters({self._store!r})'

    def get_dict_view(self) -> types.MappingProxyType:
        return types.MappingProxyType(self._store)

    def with_defaults(self, defaults: Mapping[str, Any], drop_keys_not_in_defaults: bool = False) -> Tuple['Hyperparameters', List[str]]:
        """ Returns a copy completed with `defaults` and the names of required (`...`-valued) entries that are absent. """
        defaults = dict(defaults)
        store = {n: v for n, v in self._store.items() if n in defaults} if drop_keys_not_in_defaults else dict(self._store)
        for name, value in defaults.items():
            if name not in store and value is not ...:
                store[name] = value
        missing = [n for n in defaults if n not in store]
        return Hyperparameters(**store), missing


HYPERPARAMS_T = Union[Hyperparameters, Dict[str, Any]]


def to_hyperparameters(hp: HYPERPARAMS_T, defaults: HYPERPARAMS_T = None, raise_if_missing: bool = True, drop_keys_not_in_defaults: bool = False):
    """ Dict -> `Hyperparameters`; with `defaults` returns `(hp, missing)` and raises `ValueError` on missing required entries. """
    if not isinstance(hp, Hyperparameters):
        hp = Hyperparameters(**hp)
    if defaults is None:
        return hp
    hp, missing = hp.with_defaults(defaults, drop_keys_not_in_defaults=drop_keys_not_in_defaults)
    if missing:
        msg = f'Error: Missing mandatory (hyper)parameter(s) (missing="{missing}").'
        logging.error(msg)
        if raise_if_missing:
            raise ValueError(msg)
    return hp, missing
