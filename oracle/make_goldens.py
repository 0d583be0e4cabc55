""" Generates the committed fixtures under tests/golden/ — run HERE (the build container), where /root/reference exists:

    python oracle/make_goldens.py

Test infrastructure only. Three groups of fixtures:
  1. `ref_pure_functions.json` — outputs of the handful of reference functions that CAN execute in this image. The reference package
     itself does not import, so each function is lifted from its source file with `ast` and exec'd in isolation with the
     stdlib names it uses: `get_padding_from_kernel` (meta/nn.py:393-399), `get_by_identifier` (utils.py:365-379),
     `Hyperparameters.with_defaults` (meta/data/training_metadata.py:108-118) and `to_hyperparameters` (meta/hyperparams.py:229-248).
  2. `torchvision_preprocess.pt` — the third-party arithmetic the reference's recipe delegates to, run for real: `ToTensor()` +
     `Normalize(mean, std)` from torchvision on seeded uint8 PIL images (conf/base/parameters.yml:197-201), `RandomCrop(padding, fill=0)`
     geometry via `torchvision.transforms.functional.pad/crop` and `hflip`.
  3. `oracle_default_net.pt` / `oracle_resnet_style.pt` — the oracle's own logits / loss / gradients / running statistics for fixed seeds
     (small batch), so that the GPU box (no /root/reference, maybe a different CPU) checks against the same numbers; and
     `index_maps.json` — sha256 of the integer source-index map of every (flip, top, left) at 32x32 with pad 4.
"""
import ast
import collections.abc
import copy
import hashlib
import importlib
import json
import logging
import math
import re
import sys
import types
from pathlib import Path
from typing import Any, Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path('/root/reference')
GOLDEN = ROOT / 'tests' / 'golden'


def _lift(path: Path, name: str, env: Dict[str, Any], cls: str = None):
    """ exec one top-level function (or one class) of a reference source file in `env`, without importing the file. """
    tree = ast.parse(path.read_text())
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == (cls or name):
            if isinstance(node, ast.FunctionDef):
                node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, str(path), 'exec'), env)
            return env[cls or name]
    raise KeyError(name)


def reference_pure_functions() -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    env = dict(math=math, Sequence=Sequence, logging=logging, SIZE_N_T=Any)
    gp = _lift(REF / 'src/deepcv/meta/nn.py', 'get_padding_from_kernel', env)
    # NOTE the reference returns `padding[0]` for sequence inputs (its `is_sequence` flag is inverted): an int that torch.nn.Conv2d
    # broadcasts to every dim — identical to the per-dim list for the square kernels the shipped specs use.
    out['get_padding_from_kernel'] = {str(k): gp(k) for k in ([1, 1], [3, 3], [5, 5], [7, 7], [2, 2], [4, 4], [9, 9])}

    env = dict(re=re, importlib=importlib)
    gbi = _lift(REF / 'src/deepcv/utils.py', 'get_by_identifier', env)
    out['get_by_identifier'] = {ident: f'{gbi(ident).__module__}.{gbi(ident).__qualname__}' for ident in ('torch.nn.ReLU', 'torch.nn.Flatten', 'torch.optim.AdamW', 'math.floor')}
    bad = {}
    for ident in ('torch.nn.', 'not an identifier', 'surely_not_defined_anywhere'):
        try:
            gbi(ident)
            bad[ident] = None
        except Exception as e:
            bad[ident] = type(e).__name__
    out['get_by_identifier_errors'] = bad

    class TrainingMetaData:  # stand-in base (uuid bookkeeping only: training_metadata.py:30-50)
        def __init__(self, existing_uuid=None):
            pass
    env = dict(collections=collections, types=types, Union=Union, Dict=Dict, Any=Any, Tuple=Tuple, List=List, TrainingMetaData=TrainingMetaData, uuid=types.SimpleNamespace(UUID=object))
    HP = _lift(REF / 'src/deepcv/meta/data/training_metadata.py', None, env, cls='Hyperparameters')
    env2 = dict(Hyperparameters=HP, logging=logging, Union=Union, Tuple=Tuple, List=List, HYPERPARAMS_T=Any)
    to_hp = _lift(REF / 'src/deepcv/meta/hyperparams.py', 'to_hyperparameters', env2)
    cases = []
    for hp, defaults, drop in (({'a': 1, 'b': 2}, {'a': ..., 'c': 3}, False), ({'a': 1, 'b': 2}, {'a': ..., 'c': 3}, True), ({'b': 2}, {'a': ..., 'c': None}, False),
                               ({'architecture': [], 'act_fn': 'x'}, {'architecture': ..., 'act_fn': ..., 'weight_norm': None, 'spectral_norm': None}, False)):
        enc_defaults = {k: ('...' if v is ... else v) for k, v in defaults.items()}
        try:
            res, missing = to_hp(hp, defaults, raise_if_missing=False, drop_keys_not_in_defaults=drop)
            cases.append(dict(hp=hp, defaults=enc_defaults, drop=drop, result=dict(res), missing=missing))
        except Exception as e:
            cases.append(dict(hp=hp, defaults=enc_defaults, drop=drop, error=type(e).__name__))
    out['to_hyperparameters'] = cases
    return out


def torchvision_preprocess() -> Dict[str, torch.Tensor]:
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    from PIL import Image
    g = torch.Generator().manual_seed(434546)
    out = {}
    for name, (n, s, mean, std, pad) in {'cifar': (6, 32, [0.491, 0.482, 0.447], [0.247, 0.243, 0.261], 4), 'imagenet': (2, 64, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], 8)}.items():
        imgs = torch.randint(0, 256, (n, s, s, 3), generator=g, dtype=torch.uint8)
        flip = (torch.rand(n, generator=g) < 0.5).to(torch.uint8)
        crop = torch.randint(0, 2 * pad + 1, (n, 2), generator=g, dtype=torch.int32)
        plain, aug = [], []
        tf = T.Compose([T.ToTensor(), T.Normalize(mean, std)])
        for k in range(n):
            pil = Image.fromarray(imgs[k].numpy())
            plain.append(tf(pil))
            a = TF.crop(TF.pad(pil, pad, fill=0), int(crop[k, 0]), int(crop[k, 1]), s, s)   # RandomCrop(size, padding=pad, fill=0) geometry
            a = TF.hflip(a) if int(flip[k]) else a
            aug.append(tf(a))
        out[name] = dict(images=imgs, flip=flip, crop=crop, pad=pad, mean=mean, std=std, plain=torch.stack(plain), augmented=torch.stack(aug))
    # every byte value through ToTensor+Normalize (the full value table of the transform)
    ramp = torch.arange(256, dtype=torch.uint8).view(16, 16, 1).repeat(1, 1, 3)
    out['ramp'] = dict(images=ramp[None], plain=T.Compose([T.ToTensor(), T.Normalize([0.491, 0.482, 0.447], [0.247, 0.243, 0.261])])(Image.fromarray(ramp.numpy()))[None])
    return out


def oracle_net(model_yaml: str, model_name: str, input_shape, batch: int, classes: int, seed: int = 563454) -> Dict[str, Any]:
    from deepcv_b200.yaml_config import benchmark_model_spec
    from oracle.deepcv_oracle import OracleDeepcvModule, train_step
    hp = benchmark_model_spec(ROOT / 'conf' / 'base' / model_yaml, model_name, out_features=classes)
    torch.manual_seed(seed)
    model = OracleDeepcvModule(input_shape, hp)
    # non-trivial affine parameters everywhere (GroupNorm / biases start at 1 / 0): parity must not hide behind zeros
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    init_state = copy.deepcopy(model.state_dict())
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, *input_shape, generator=g)
    y = torch.randint(0, classes, (batch,), generator=g)
    loss, logits = train_step(model, x, y)
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    # the same step in float64: the arbiter for gradients that are analytically (near) zero (SURVEY.md section 8.d parity gates)
    model64 = OracleDeepcvModule(input_shape, hp)
    model64.load_state_dict(init_state)
    model64 = model64.double()
    loss64, logits64 = train_step(model64, x.double(), y)
    grads64 = {n: p.grad.clone() for n, p in model64.named_parameters()}
    return dict(loss64=loss64, logits64=logits64, grads64=grads64, hp_yaml=model_yaml, model_name=model_name, input_shape=tuple(input_shape), classes=classes, seed=seed, state=init_state, x=x, y=y, loss=loss, logits=logits,
                grads=grads, state_after=copy.deepcopy(model.state_dict()))


def index_maps() -> Dict[str, str]:
    from oracle.deepcv_oracle import preprocess_index_map
    out = {}
    for flip in (0, 1):
        for top in range(9):
            for left in range(9):
                m = preprocess_index_map(32, 32, 32, 32, 4, flip, top, left)
                out[f'{flip},{top},{left}'] = hashlib.sha256(np.ascontiguousarray(m.astype('<i4')).tobytes()).hexdigest()
    return out


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    (GOLDEN / 'ref_pure_functions.json').write_text(json.dumps(reference_pure_functions(), indent=1, sort_keys=True))
    torch.save(torchvision_preprocess(), GOLDEN / 'torchvision_preprocess.pt')
    torch.save(oracle_net('parameters.yml', 'image_classifier', (3, 32, 32), 8, 10), GOLDEN / 'oracle_default_net.pt')
    (GOLDEN / 'index_maps.json').write_text(json.dumps(index_maps(), indent=0, sort_keys=True))
    print('wrote', sorted(p.name for p in GOLDEN.iterdir()))


if __name__ == '__main__':
    main()
