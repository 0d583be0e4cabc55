""" Image classification entry points — host-side mirror of `src/deepcv/classification/image.py:40-80`.

`create_model(datasets, model_params)` and `train(datasets, model, hp)` keep the reference's names and argument meaning; the Kedro
pipeline wiring (`get_pipelines`, :28-38) is orchestration outside the hot path and is not rebuilt. Differences, all from the defect
ledger: the loss is handed to `ignite_training.train` as `losses=` (the reference passes `loss=`, a
`TypeError`); the head's `out_features` test uses `in` on the params dict (the reference uses `hasattr` on a dict, always False);
`spectral_norm` must be null (`base_module.py:109-111` raises on it). The loss and optimizer are the library's fused
cross-entropy and the flat-buffer AdamW (`torch.nn.CrossEntropyLoss()` / `torch.optim.AdamW` in the reference, :70-71).
"""
import copy
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

from ..meta import base_module, ignite_training
from ..meta.flat_params import FlatAdamW
from ..meta.hyperparams import HYPERPARAMS_T

__all__ = ['create_model', 'train']


def _recursive_classes(dataset) -> Optional[list]:
    """ `deepcv.utils.recursive_getattr(dataset, 'classes', recurse_on_type=Dataset)` (reference `utils.py`): looks through dataset wrappers. """
    seen = set()
    while dataset is not None and id(dataset) not in seen:
        seen.add(id(dataset))
        if getattr(dataset, 'classes', None) is not None:
            return list(dataset.classes)
        dataset = next((v for v in vars(dataset).values() if isinstance(v, Dataset)), None) if hasattr(dataset, '__dict__') else None
    return None


def create_model(datasets: Dict[str, Dataset], model_params: HYPERPARAMS_T) -> torch.nn.Module:
    """ reference :40-54: input shape from the first training sample, head width from the dataset's classes (or target shape). """
    dummy_img, dummy_target = datasets['trainset'][0]
    model_params = dict(model_params)
    model_params['architecture'] = copy.deepcopy(model_params['architecture'])
    head = model_params['architecture'][-1]['fully_connected']
    if 'out_features' not in head:
        classes = _recursive_classes(datasets['trainset'])
        if classes is not None:
            head['out_features'] = len(classes)
        elif isinstance(dummy_target, torch.Tensor):
            head['out_features'] = int(np.prod(dummy_target.shape))
    input_shape = tuple(dummy_img.shape)
    if dummy_img.dtype == torch.uint8 and len(input_shape) == 3:   # fused recipe: samples stay uint8 HWC until the batch kernel (FusedPreprocess)
        input_shape = (input_shape[2], input_shape[0], input_shape[1])
    return base_module.DeepcvModule(input_shape, model_params)


def train(datasets: Dict[str, Dataset], model: torch.nn.Module, hp: HYPERPARAMS_T) -> Tuple[Dict[str, float], Any, Optional[str]]:
    """ reference :65-84 (the non-NAS branch): `ignite_training.train` with cross-entropy and AdamW. """
    hp = dict(hp)
    backend_conf = ignite_training.BackendConfig(**(hp.pop('backend_conf') if 'backend_conf' in hp else {}))
    return (*ignite_training.train(hp=hp, model=model, losses=ignite_training.CrossEntropyLoss(), datasets=datasets, opt=FlatAdamW, backend_conf=backend_conf,
                                   metrics=None, callbacks_handler=None, nni_compression_pruner=None), None)
