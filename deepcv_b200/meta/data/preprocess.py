""" Preprocess-recipe API — host-side mirror of `src/deepcv/meta/data/preprocess.py`, plus the fused device transform.

Kept verbatim from the reference: `PreprocessedDataset` (:35-63), `TRANSFORM_ARGS_PROCESSORS` / `register_transform_processor`
(:89-106), `_process_normalization_stats` (:109-134), `_parse_transforms_specification` (:137-178), `split_dataset` (:181-215) and
`preprocess(params, dataset_or_trainset, testset)` (:218-260). The recipe (`transforms:` list of types or `{type: kwargs}`) is
parsed exactly the same way, so the reference's `cifar10_preprocessing` recipe still builds a torchvision `Compose`.

New: `FusedPreprocess`, a YAML-nameable transform type. Per sample (DataLoader worker side) it only makes sure the image is a uint8
H x W x C tensor; per batch (device side) it is the single HBM-bound kernel `dcv_preprocess_u8`: pad-crop -> horizontal flip on the
uint8 image, then ToTensor (/255) and Normalize ((v - mean)/std) — the op order of torchvision's ToTensor + Normalize
(`conf/base/parameters.yml:197-210`) with the flip / crop geometry the augmentation recipe names (`augmentation.py:39-44`,
`parameters.yml:152,157`). It is a `torch.nn.Module`, so `torch.nn.Sequential(FusedPreprocess(...), model)` leaves the training step
(`y_pred = model(x)`, ignite_training.py:248) untouched.
"""
import functools
import logging
from typing import Any, Callable, Dict, Iterable, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from ... import ops, utils
from ..hyperparams import Hyperparameters, to_hyperparameters

__all__ = ['PreprocessedDataset', 'fn_to_transform', 'TRANSFORM_ARGS_PROCESSORS', 'register_transform_processor', 'split_dataset', 'preprocess', 'FusedPreprocess',
           'draw_augmentation_params']
NL = '\n'


class PreprocessedDataset(Dataset):
    """ Applies input / target transforms to the items of an underlying dataset (reference :35-63). """

    def __init__(self, underlying_dataset: Dataset, img_transform: Optional[Callable], target_transform: Callable = None, augmentation_transform: Callable = None):
        self._underlying_dataset = underlying_dataset
        self._img_transform = img_transform
        self._target_transform = target_transform
        self._augmentation_transform = augmentation_transform

    def __getitem__(self, index):
        data = self._underlying_dataset.__getitem__(index)
        if isinstance(data, tuple):
            x, *ys = data
            if self._img_transform is not None:
                x = self._img_transform(x)
            if self._target_transform is not None:
                ys = (self._target_transform(y) for y in ys)
            if self._augmentation_transform is not None:
                raise NotImplementedError
            return (x, *ys)
        return self._img_transform(data)

    def __len__(self):
        return len(self._underlying_dataset)

    def __repr__(self):
        return f'{PreprocessedDataset.__name__}[{repr(vars(self))}]'


def fn_to_transform(fn: Callable, *transform_args: Iterable[str]) -> Callable[[], Callable]:
    def _get_transform(**transform_kwargs) -> Callable:
        if transform_kwargs:
            if not set(transform_args).issuperset(transform_kwargs.keys()):
                raise ValueError(f'Error: `{fn}` transform expected following arguments: `{transform_args}` but got: `{transform_kwargs}`')
            return functools.partial(fn, **transform_kwargs)
        return fn
    return _get_transform


""" Maps transform types to (arguments-processing function, names of the arguments it can compute at runtime from the trainset). """
TRANSFORM_ARGS_PROCESSORS: Dict[Any, Tuple[Callable, Iterable[str]]] = dict()


def register_transform_processor(transform: Union[str, Callable], processable_args_names: Iterable[str]):
    """ Decorator registering `process_fn(trainset=, to_process=) -> dict` for `transform` (reference :92-106). """
    def _wrap(process_fn: Callable[..., Dict[str, Any]]):
        if transform in TRANSFORM_ARGS_PROCESSORS:
            raise RuntimeError(f'Error: {transform} is already registered in `deepcv.meta.data.preprocess.TRANSFORM_ARGS_PROCESSORS`')
        TRANSFORM_ARGS_PROCESSORS[transform] = (process_fn, processable_args_names)
        return process_fn
    return _wrap


def _to_float_chw(img) -> torch.Tensor:
    """ torchvision `ToTensor` semantics for the statistics pass: HWC uint8 (PIL / numpy / tensor) -> CHW float32 in [0, 1]. """
    if isinstance(img, torch.Tensor):
        if img.dtype == torch.uint8:
            return img.permute(2, 0, 1).to(torch.float32).div(255) if img.dim() == 3 else img.to(torch.float32).div(255)
        return img
    arr = np.asarray(img)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1)
    return t.to(torch.float32).div(255) if t.dtype == torch.uint8 else t.to(torch.float32)


def _process_normalization_stats(trainset: Dataset, to_process: Sequence[str]) -> Dict[str, torch.Tensor]:
    """ Per-channel mean of per-image means and mean of per-image stds over the trainset (reference :109-134). """
    assert {'mean', 'std'}.issuperset(to_process), f'Error: `_process_normalization_stats` can only process `mean` or `std`, not: `{to_process}`'
    first = trainset[0][0] if isinstance(trainset[0], tuple) else trainset[0]
    channels = _to_float_chw(first).shape[0]
    stats = {n: torch.zeros((channels,)) for n in to_process}
    for input_data in trainset:
        img = _to_float_chw(input_data[0] if isinstance(input_data, tuple) else input_data)
        dims = tuple(range(1, len(img.shape)))
        if 'mean' in stats:
            stats['mean'] += img.mean(dim=dims) / len(trainset)
        if 'std' in stats:
            stats['std'] += img.std(dim=dims) / len(trainset)
    return stats


def _parse_transforms_specification(transform_identifiers: Sequence, trainset: Dataset, transform_args_processors: Dict = TRANSFORM_ARGS_PROCESSORS):
    """ Recipe -> composed transform (reference :137-178): each entry is a callable / type, a dotted string identifier, or a
    single-key `{type: kwargs}` mapping; arguments missing from the YAML are computed by the registered processor, if any. """
    fn_name = '_parse_transforms_specification'
    transforms = []
    for spec in transform_identifiers:
        transform_kwargs = {}
        if isinstance(spec, dict):
            if not len(spec.items()) == 1:
                raise ValueError(f'Error: {fn_name}: Invalid transform specification, a transform should be specified by a single transform '
                                 f'type/identifer which can eventually be mapped to a dict of keyword arguments')
            if not isinstance(next(iter(spec.values())), dict):
                raise ValueError(f'Error: {fn_name}: A value mapped to a transform is expected to be a dict of keyword arguments which will '
                                 f'be provided to transform\'s constructor/function, got: `{spec}`')
            spec, transform_kwargs = next(iter(spec.items()))
            transform_kwargs = dict(transform_kwargs)
        if isinstance(spec, str):
            spec = utils.get_by_identifier(spec)
        elif not callable(spec):
            raise ValueError(f'Error: {fn_name} couldn\'t find `{spec}` tranform, transform specification should either be a string identifier or tranform `Callable` type.')

        if spec in transform_args_processors:
            process_fn, processable_args_names = transform_args_processors[spec]
            to_process = [arg_name for arg_name in processable_args_names if arg_name not in transform_kwargs]
            if len(to_process) > 0:
                processed_state = process_fn(trainset=trainset, to_process=to_process)
                transform_kwargs.update({n: processed_state[n] for n in to_process})
        transforms.append(spec(**transform_kwargs))
    return _compose(transforms)


def _compose(transforms):
    try:
        import torchvision
        return torchvision.transforms.Compose(transforms)
    except ImportError:  # torchvision is optional for the fused recipe
        def _composed(x):
            for t in transforms:
                x = t(x)
            return x
        _composed.transforms = transforms
        return _composed


def split_dataset(params, dataset_or_trainset: Dataset, testset: Dataset = None) -> Dict[str, Dataset]:
    """ trainset / validset / testset split by ratios (reference :181-215). """
    params, _ = to_hyperparameters(params, defaults={'validset_ratio': None, 'testset_ratio': None, 'cache': False})
    testset_ratio, validset_ratio = params['testset_ratio'], params['validset_ratio']
    split_lengths = tuple()
    if testset is None:
        if testset_ratio is None:
            raise ValueError(f'Error: split_dataset function either needs an existing `testset` as argument or you must specify a `testset_ratio` in `params` '
                             f'(probably from parameters/preprocessing YAML config){NL}Provided dataset spliting parameters: "{params}"')
        split_lengths += (int(len(dataset_or_trainset) * testset_ratio),)
    if validset_ratio is not None:
        split_lengths += (int(len(dataset_or_trainset) * validset_ratio),)
    elif testset is not None:
        return {'trainset': dataset_or_trainset, 'testset': testset}
    trainset_size = int(len(dataset_or_trainset) - np.sum(split_lengths))
    if trainset_size < 1:
        raise RuntimeError(f'Error in split_dataset: testset and eventual validset size(s) are too large, there is no remaining trainset samples{NL}'
                           f'(`len(dataset_or_trainset)={len(dataset_or_trainset)}`, `testset_ratio={testset_ratio}`, `validset_ratio={validset_ratio}`)')
    trainset, *testset_and_validset = torch.utils.data.random_split(dataset_or_trainset, (trainset_size, *split_lengths))
    if testset is None:
        testset = testset_and_validset[0]
    validset = testset_and_validset[-1] if validset_ratio is not None else None
    if params['cache']:
        raise NotImplementedError
    return {'trainset': trainset, 'validset': validset, 'testset': testset} if validset else {'trainset': trainset, 'testset': testset}


def preprocess(params, dataset_or_trainset: Dataset, testset: Optional[Dataset]) -> Dict[str, PreprocessedDataset]:
    """ Main preprocessing procedure (reference :218-260): seed, split, parse the `transforms` recipe, wrap datasets. """
    params, _ = to_hyperparameters(params, defaults={'transforms': ..., 'target_transforms': [], 'cache': False, 'augmentation_reciepe': None, 'split_dataset': {}, 'seed': None})
    if params['seed'] is not None:
        utils.set_seeds(params['seed'])
    datasets = split_dataset(params['split_dataset'], dataset_or_trainset, testset)
    preprocess_transforms = dict(img_transform=_parse_transforms_specification(params['transforms'], trainset=datasets['trainset']))
    if params['target_transforms'] is not None and len(params['target_transforms']) > 0:
        preprocess_transforms['target_transform'] = _parse_transforms_specification(params['target_transforms'], trainset=datasets['trainset'])
    if params['augmentation_reciepe'] is not None:
        raise NotImplementedError('Error: augmentation recipes are unimplemented in the reference too (`apply_augmentation_reciepe` raises); use `FusedPreprocess(pad=..., flip=...)`')
    datasets = {n: PreprocessedDataset(ds, **preprocess_transforms) for n, ds in datasets.items()}
    if params['cache']:
        raise NotImplementedError
    return datasets


# --------------------------------------------------------------------------------------------------------------------------------
# Fused device-side transform

def draw_augmentation_params(n: int, pad: int, generator: torch.Generator, flip: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """ Host-side draw of the per-sample augmentation parameters: `flip ~ rand(n) < 0.5`, then
    `(top, left) ~ randint(0, 2*pad + 1)`. Drawn on the host so that the index selection is bit-exact by construction. """
    flips = (torch.rand(n, generator=generator) < 0.5).to(torch.uint8) if flip else torch.zeros(n, dtype=torch.uint8)
    crops = torch.randint(0, 2 * pad + 1, (n, 2), generator=generator, dtype=torch.int32)
    return flips, crops


class FusedPreprocess(torch.nn.Module):
    """ `ToTensor` + `Normalize(mean, std)` (+ random pad-crop and horizontal flip) as one uint8 -> float kernel.

    Args (YAML kwargs): `mean`, `std` per channel (computed from the trainset when absent, like `Normalize`'s, through
    `TRANSFORM_ARGS_PROCESSORS`); `pad` zero padding for the random crop (0 = no crop); `flip` random horizontal flip;
    `out_size` crop size (default: input size); `dtype` 'float32' | 'bfloat16'; `seed` of the host generator drawing flips / offsets.
    """

    def __init__(self, mean: Sequence[float], std: Sequence[float], pad: int = 0, flip: bool = False, out_size: Optional[Sequence[int]] = None,
                 dtype: Union[str, torch.dtype] = torch.float32, seed: int = 434546):
        super().__init__()
        self.register_buffer('mean', torch.as_tensor(mean, dtype=torch.float32).reshape(-1).clone())
        self.register_buffer('std', torch.as_tensor(std, dtype=torch.float32).reshape(-1).clone())
        self.pad, self.flip = int(pad), bool(flip)
        self.out_size = None if out_size is None else (int(out_size[0]), int(out_size[1]))
        self.dtype = getattr(torch, dtype) if isinstance(dtype, str) else dtype
        self.generator = torch.Generator().manual_seed(int(seed))
        self._pinned = None

    def extra_repr(self) -> str:
        return f'mean={self.mean.tolist()}, std={self.std.tolist()}, pad={self.pad}, flip={self.flip}, out_size={self.out_size}, dtype={self.dtype}'

    def draw(self, n: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """ Per-sample (flip, crop offsets) for a batch of `n`, or (None, None) outside training / when no augmentation is configured. """
        if not self.training or (self.pad == 0 and not self.flip and self.out_size is None):
            return None, None
        return draw_augmentation_params(n, self.pad, self.generator, flip=self.flip)

    def forward(self, img, flip: Optional[torch.Tensor] = None, crop_yx: Optional[torch.Tensor] = None) -> torch.Tensor:
        # ---- per-sample call (CPU, inside a DataLoader worker): keep the image as uint8 H x W x C; the arithmetic happens per batch on the device
        if not isinstance(img, torch.Tensor) or (img.dim() == 3 and not img.is_cuda):
            arr = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(img)))
            if arr.dim() == 2:
                arr = arr.unsqueeze(-1)
            if arr.dtype != torch.uint8:
                raise TypeError(f'FusedPreprocess expects uint8 images, got {arr.dtype}')
            return arr
        # ---- per-batch call (device)
        if img.dim() != 4 or img.dtype != torch.uint8:
            raise TypeError(f'FusedPreprocess expects a uint8 N x H x W x C batch, got {img.dtype} {tuple(img.shape)}')
        if not img.is_cuda:
            raise RuntimeError('deepcv_b200: FusedPreprocess batch transform runs on CUDA (sm_100a) only; move the uint8 batch to the device first')
        n, h, w, c = img.shape
        if c != self.mean.numel():
            raise ValueError(f'FusedPreprocess configured for {self.mean.numel()} channel(s), got {c}')
        if flip is None and crop_yx is None:
            flip, crop_yx = self.draw(n)
        if flip is not None and not flip.is_cuda:
            flip = flip.pin_memory().to(img.device, non_blocking=True) if torch.cuda.is_available() else flip
        if crop_yx is not None and not crop_yx.is_cuda:
            crop_yx = crop_yx.pin_memory().to(img.device, non_blocking=True) if torch.cuda.is_available() else crop_yx
        out_hw = self.out_size if self.out_size is not None else (h, w)
        return ops.preprocess_u8(img.contiguous(), self.mean, self.std, flip=flip, crop_yx=crop_yx, pad=self.pad, out_hw=out_hw, dtype=self.dtype)


register_transform_processor(transform=FusedPreprocess, processable_args_names=['mean', 'std'])(_process_normalization_stats)
try:  # the reference registers the same processor for torchvision's Normalize (:108)
    import torchvision as _tv
    register_transform_processor(transform=_tv.transforms.Normalize, processable_args_names=['mean', 'std'])(_process_normalization_stats)
except ImportError:
    pass
