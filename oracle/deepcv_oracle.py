""" CPU ORACLE — test infrastructure only, never the product path.

A working restatement, from stock `torch.nn` CPU modules, of what the reference's DeepcvModule hot path is *meant* to
compute. The reference cannot be imported in this image (Python 3.12 removed `imp`; kedro / ignite /
nni / mlflow / albumentations are absent; three files do not parse — SURVEY.md §8.c.2), and its own tests assert nothing
about tensors, so: **parity unpinned by the reference's tests**. What pins this oracle instead:
  * its arithmetic is performed by the very third-party kernels the reference delegates to (torch.nn.Conv2d,
    BatchNorm2d, GroupNorm, AvgPool2d, F.interpolate, Linear; torchvision ToTensor/Normalize semantics),
  * `oracle/make_goldens.py` executes the handful of pure functions of the reference that *can* run here
    (`get_padding_from_kernel`, `Hyperparameters.with_defaults`, `get_by_identifier`) and freezes their outputs and the
    oracle's own outputs as fixtures under `tests/golden/`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` arm may import this module.

Each function cites the reference lines it follows (paths relative to /root/reference/src/deepcv/).
"""
import copy
import inspect
import math
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------------
# Spec tokens — meta/nn_spec.py:35-50 (compared on their string values; `_nested_deepcvmodule` is the spelling the
# YAML actually uses, conf/base/parameters.yml:85)
FROM, SUBMODULE_NAME = '_from', '_name'
NESTED = ('_nested_deepcv_module', '_nested_deepcvmodule')
NEW_BRANCH = '_new_branch_from_tensor'
HP_DEFAULTS = {'architecture': ..., 'act_fn': ..., 'weight_norm': None, 'spectral_norm': None}  # meta/base_module.py:68


def get_padding_from_kernel(kernel_size):
    """ meta/nn.py:393-399: floor((k-1)/2) per dim. """
    seq = isinstance(kernel_size, Sequence)
    pad = [max(0, math.floor((k - 1.) / 2.)) for k in (kernel_size if seq else [kernel_size])]
    return pad if seq else pad[0]


def normalization_ops(shape: Sequence[int], batch_norm=None, layer_norm=None, instance_norm=None, group_norm=None) -> List[torch.nn.Module]:
    """ meta/nn.py:448-516: BN -> LN -> IN -> GN, only the configured ones, feature counts from `shape` (C, *spatial). """
    ops, dims = [], len(shape) - 1
    if batch_norm:
        ops.append({1: torch.nn.BatchNorm1d, 2: torch.nn.BatchNorm2d, 3: torch.nn.BatchNorm3d}[max(dims, 1)](num_features=shape[0], **batch_norm))
    if layer_norm:
        ops.append(torch.nn.LayerNorm(normalized_shape=list(shape[1:]), **layer_norm))
    if instance_norm:
        ops.append({1: torch.nn.InstanceNorm1d, 2: torch.nn.InstanceNorm2d, 3: torch.nn.InstanceNorm3d}[dims](num_features=shape[0], **instance_norm))
    if group_norm:
        ops.append(torch.nn.GroupNorm(num_channels=shape[0], **group_norm))
    return ops


def layer(layer_op: torch.nn.Module, act_fn, dropout_prob=None, preactivation=False, input_shape=None, **norms) -> torch.nn.Sequential:
    """ meta/nn.py:519-554: post-act `[Dropout] -> op -> act -> norms`; pre-act `[Dropout] -> norms -> act -> op`. """
    norm_ops = []
    if any(norms.values()):
        shape = list(input_shape)
        if not preactivation:  # normalised tensor is the op's output: meta/nn.py:545-548
            with torch.no_grad():
                shape = list(layer_op(torch.zeros(1, *input_shape)).shape[1:])
        norm_ops = normalization_ops(shape, **{k: copy.deepcopy(v) for k, v in norms.items()})
    drop = torch.nn.Dropout(p=dropout_prob) if dropout_prob not in (None, 0., 0) else None
    act = act_fn() if act_fn is not None else None
    ops = (drop, *norm_ops, act, layer_op) if preactivation else (drop, layer_op, act, *norm_ops)
    return torch.nn.Sequential(*(m for m in ops if m is not None))


class Link(torch.nn.Module):
    """ meta/submodule_creators.py:272-332 (+ reduction fns :43-65 as intended): out=[x]+refs, each ref
    bilinearly rescaled to x's spatial shape iff shapes differ and `allow_scaling`; then sum / mean / concat(dim=1). """

    def __init__(self, reduction: str, allow_scaling: bool, align_corners: bool = False, ignore_input: bool = False, apply_in_parallel: bool = True):
        super().__init__()
        self.reduction, self.allow_scaling, self.align_corners, self.ignore_input = reduction, allow_scaling, align_corners, ignore_input
        self.apply_in_parallel = apply_in_parallel
        self.referenced_submodules: List[str] = []

    def forward(self, x, referenced_submodules_out: 'OrderedDict[str, torch.Tensor]'):
        """ meta/nn.py:130-194 call convention: operands are lists of tensors (parallel branches); applied in parallel, the i-th input is reduced with
        the i-th tensor of every reference (all must hold as many tensors); otherwise everything is reduced together. Branch order is kept (the
        reference's `list.pop()` would reverse it at every parallel submodule: SURVEY.md section 8.c.2-style defect, not reproduced). """
        xs = [] if self.ignore_input else (list(x) if isinstance(x, (list, tuple)) else [x])
        refs = [referenced_submodules_out[name] for name in self.referenced_submodules]
        refs = [list(r) if isinstance(r, (list, tuple)) else [r] for r in refs]
        if self.apply_in_parallel:
            if not self.ignore_input and not all(len(r) == len(xs) for r in refs):
                raise ValueError('Error: When `in_tensors_count_similar_to_refs` is `True`, all referenced output tensor(s) should each have as many tensor(s) as input tensor(s)')
            if not all(len(r) == len(refs[0]) for r in refs):
                raise ValueError('Error: When `refs_tensor_count_similar` is `True`, all referenced output tensor(s) should each have as many tensor(s)')
            outs = []
            for i in range(len(refs[0])):
                o = self._reduce(([] if self.ignore_input else [xs[i]]), [r[i] for r in refs])
                outs.extend(o if isinstance(o, list) else [o])
        else:
            o = self._reduce(xs, [t for r in refs for t in r])
            outs = o if isinstance(o, list) else [o]
        return outs[0] if len(outs) == 1 else outs

    def _reduce(self, out: List[torch.Tensor], ref_tensors: List[torch.Tensor]):
        out = list(out)
        for y in ref_tensors:
            target = (out[0] if out else y).shape[2:]
            if y.shape[2:] != target:
                if not self.allow_scaling:
                    raise RuntimeError(f"Error: Couldn't forward throught link: residual_shape='{y.shape}' != prev_features_shape='{out[0].shape}'")
                mode = {1: 'linear', 2: 'bilinear', 3: 'trilinear'}[len(target)]
                y = F.interpolate(y, size=target, mode=mode, align_corners=self.align_corners)  # meta/nn.py:665-676
            out.append(y)
        if self.reduction == 'none' or len(out) == 1:
            return out if len(out) > 1 else out[0]
        if self.reduction == 'concat':
            return torch.cat(out, dim=1)
        if self.reduction == 'sum':
            return torch.stack(out, 0).sum(0) if len(out) > 1 else out[0]
        if self.reduction == 'mean':
            return torch.stack(out, 0).mean(0) if len(out) > 1 else out[0]
        raise ValueError(f'Error: Invalid "{self.reduction}" reduction function name.')


class Reduce(torch.nn.Module):
    """ meta/submodule_creators.py:179-186 (`reduce`) and :188-200 (`select_tensor`): reduction / selection over the parallel tensors of the previous
    submodule (not applied per tensor: `takes_list`). """
    takes_list = True

    def __init__(self, reduction: str, pick=None):
        super().__init__()
        self.reduction, self.pick = reduction, pick

    def forward(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        if self.pick is not None:
            xs = xs[self.pick]
            xs = xs if isinstance(xs, list) else [xs]
        if len(xs) == 1 or self.reduction == 'none':
            return xs[0] if len(xs) == 1 else xs
        if self.reduction == 'concat':
            return torch.cat(xs, dim=1)
        return torch.stack(xs, 0).sum(0) if self.reduction == 'sum' else torch.stack(xs, 0).mean(0)


def _parse_slice(spec):
    if isinstance(spec, int):
        return spec
    parts = [int(v) if v.strip() else None for v in str(spec).split(':')]
    return parts[0] if len(parts) == 1 else slice(*parts)


def _conv_or_linear_creator(op_t):
    def creator(submodule_params, input_shape, act_fn=None, dropout_prob=None, preactivation=False, batch_norm=None, layer_norm=None, instance_norm=None, group_norm=None):
        """ meta/submodule_creators.py:237-255 """
        p = dict(submodule_params)
        if issubclass(op_t, torch.nn.Linear):
            p.setdefault('in_features', int(np.prod(input_shape)))
        else:
            if 'padding' not in p:
                p['padding'] = get_padding_from_kernel(p['kernel_size'])
            p.setdefault('in_channels', input_shape[0])
        return layer(op_t(**p), act_fn=act_fn, dropout_prob=dropout_prob, preactivation=preactivation, input_shape=input_shape,
                     batch_norm=batch_norm, layer_norm=layer_norm, instance_norm=instance_norm, group_norm=group_norm)
    return creator


def _avg_pooling(submodule_params, input_shape):
    """ meta/submodule_creators.py:163-176 -> meta/nn.py:416 """
    return {1: torch.nn.AvgPool1d, 2: torch.nn.AvgPool2d, 3: torch.nn.AvgPool3d}[len(input_shape) - 1](**submodule_params)


def _link(is_residual):
    def creator(submodule_params, allow_scaling=False, scaling_align_corners=False, reduction='sum' if is_residual else 'concat', apply_in_parallel=True):
        if FROM not in submodule_params:
            raise ValueError('Error: Missing "_from" parameter in link YAML specification')
        return Link(reduction, allow_scaling, scaling_align_corners, apply_in_parallel=apply_in_parallel)
    return creator


def _new_branch(submodule_params, reduction='concat'):
    """ meta/submodule_creators.py:203-224: ignores the previous output, reduces the referenced tensors. """
    return Link(reduction, allow_scaling=False, ignore_input=True)


CREATORS: Dict[str, Callable] = {
    'conv1d': _conv_or_linear_creator(torch.nn.Conv1d), 'conv2d': _conv_or_linear_creator(torch.nn.Conv2d), 'conv3d': _conv_or_linear_creator(torch.nn.Conv3d),
    'linear': _conv_or_linear_creator(torch.nn.Linear), 'fully_connected': _conv_or_linear_creator(torch.nn.Linear),
    'average_pooling': _avg_pooling, 'avg_pooling': _avg_pooling,  # both spellings: SURVEY.md §8.c.2
    'residual_link': _link(True), 'dense_link': _link(False), NEW_BRANCH: _new_branch,
    'reduce': lambda submodule_params, fn, keep_dim=False: Reduce(fn),
    'select_tensor': lambda submodule_params, reduction='none': Reduce(reduction, pick=_parse_slice(submodule_params['slice'])),
}


class OracleDeepcvModule(torch.nn.Module):
    """ meta/base_module.py:39-264 + meta/nn_spec.py:55-243, with the defect ledger of SURVEY.md §8.c.2 applied. """

    def __init__(self, input_shape, hp, creators: Dict[str, Callable] = None):
        super().__init__()
        hp = dict(hp)
        missing = [k for k, v in HP_DEFAULTS.items() if v is ... and k not in hp]
        if missing:
            raise ValueError(f'Error: Missing mandatory (hyper)parameter(s) (missing="{missing}").')
        self._hp = {**{k: v for k, v in HP_DEFAULTS.items() if v is not ...}, **hp}
        self._input_shape = tuple(input_shape)
        self._creators = dict(CREATORS if creators is None else creators)
        self._features_shapes = [self._input_shape]
        self._submodules: 'OrderedDict[str, torch.nn.Module]' = OrderedDict()
        self._submodule_references: Dict[str, List[str]] = {}
        for i, spec in enumerate(self._hp['architecture']):
            name, module = self._parse(spec, i)
            self._submodules[name] = module
            refs = getattr(module, 'referenced_submodules', None)
            if refs:
                bad = [r for r in refs if r not in self._submodules]
                if bad:
                    raise ValueError(f'Error: Invalid sub-module reference(s), cant find following sub-module name(s)/label(s): "{bad}".')
                self._submodule_references[name] = list(refs)
            self._child_modules = torch.nn.ModuleDict(self._submodules)
            self._features_shapes.append(self._infer_shape())
        self._initialize_parameters(self._hp['act_fn'])

    # nn_spec.py:194-215
    @staticmethod
    def _name_params_type(spec, default_name, existing):
        subm_type, params = (list(spec.items())[0] if isinstance(spec, dict) else (spec, {}))
        name = default_name
        if isinstance(params, (list, tuple)):
            name, params = params[0], params[1]
        elif isinstance(params, str):
            name, params = params, {}
        elif isinstance(params, dict) and SUBMODULE_NAME in params:
            params = dict(params)
            name = params.pop(SUBMODULE_NAME)
        if name in existing or name == '' or not isinstance(name, str):
            raise ValueError(f'Error: Invalid or duplicate sub-module name/label: "{name}"')
        if not isinstance(params, dict):
            raise RuntimeError(f'Error: Architecture sub-module spec. must either be a parameters Dict, or a submodule name along with parameters Dict, but got: "{params}".')
        return name, dict(params), subm_type

    # nn_spec.py:107-191
    def _parse(self, spec, pos):
        name, params, subm_type = self._name_params_type(spec, f'_submodule_{pos}', self._submodules.keys())
        with_globals = {n: copy.deepcopy(v) for n, v in self._hp.items() if n not in params}
        with_globals.update(params)
        if isinstance(subm_type, str) and subm_type in NESTED:
            return name, OracleDeepcvModule(self._features_shapes[-1], with_globals, self._creators)
        fn = self._creators.get(subm_type) if isinstance(subm_type, str) else subm_type
        if fn is None:
            import importlib
            mod, _, attr = subm_type.rpartition('.')
            fn = getattr(importlib.import_module(mod), attr)
        sig = inspect.signature(fn).parameters
        last = self._features_shapes[-1]
        last = last if isinstance(last[0], (int, np.integer)) else last[0]   # parallel tensors share the layer: sized for the first (all alike)
        with_globals.update(prev_shapes=self._features_shapes, input_shape=last, input_shapes=self._features_shapes[-1])
        provided = {n: v for n, v in with_globals.items() if n in sig}
        if 'submodule_params' in sig:
            provided['submodule_params'] = {n: v for n, v in params.items() if n not in provided}
        module = fn(**provided)
        if not isinstance(module, torch.nn.Module):
            raise RuntimeError('Error: Invalid sub-module creator function or type')
        if isinstance(module, Link):
            if FROM not in params:
                raise ValueError('Error: link submodules need a "_from" parameter')
            module.referenced_submodules = [params[FROM]] if isinstance(params[FROM], str) else list(params[FROM])
        return name, module

    def _infer_shape(self):
        was_training = self.training
        self.eval()
        with torch.no_grad():
            out = self(torch.zeros(1, *self._input_shape))
        self.train(was_training)
        return tuple(out.shape[1:]) if isinstance(out, torch.Tensor) else [tuple(t.shape[1:]) for t in out]

    # base_module.py:113-155 (return after the loop; refs released after the last referrer)
    def forward(self, x):
        kept: Dict[str, torch.Tensor] = {}
        remaining = dict(self._submodule_references)
        for name, subm in self._child_modules.items():
            refs = getattr(subm, 'referenced_submodules', None)
            if refs:
                current = OrderedDict((r, kept[r]) for r in refs)
                del remaining[name]
                for r in refs:
                    if not any(r in v for v in remaining.values()):
                        kept.pop(r, None)
                x = subm(x, referenced_submodules_out=current)
            elif isinstance(x, (list, tuple)) and not getattr(subm, 'takes_list', False):   # meta/submodule_creators.py:175,254: layers / poolings are applied to each parallel tensor (shared weights)
                x = [subm(t) for t in x]
                x = x[0] if len(x) == 1 else x
            else:
                x = subm(x)
            if any(name in v for v in remaining.values()):
                kept[name] = x
        return x

    # base_module.py:230-264 (+ nn.py:585-605 gain lookup)
    def _initialize_parameters(self, act_fn):
        gain_name = {torch.nn.ReLU: 'relu', torch.nn.LeakyReLU: 'leaky_relu', torch.nn.Tanh: 'tanh', torch.nn.Sigmoid: 'sigmoid', torch.nn.Identity: 'linear'}.get(act_fn, 'relu')
        gain = torch.nn.init.calculate_gain(gain_name) if act_fn else None

        def _init(m):
            if isinstance(m, torch.nn.modules.conv._ConvNd):
                torch.nn.init.xavier_normal_(m.weight.data, gain=gain)
                if m.bias is not None:
                    m.bias.data.fill_(0.)
            elif isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight.data, gain=gain)
                if m.bias is not None:
                    m.bias.data.fill_(0.)
            elif type(m).__module__ == torch.nn.BatchNorm2d.__module__:
                torch.nn.init.uniform_(m.weight.data)
                m.bias.data.fill_(0.)
        self.apply(_init)


# --------------------------------------------------------------------------------------------------------------------
# Preprocess / augmentation: crop -> flip on uint8, then ToTensor (/255) and Normalize.

def draw_augmentation_params(n: int, pad: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """ Host-side draw shared by oracle and kernel: flip ~ rand<0.5, (top,left) ~ randint(0, 2*pad+1). """
    g = torch.Generator().manual_seed(seed)
    flip = (torch.rand(n, generator=g) < 0.5).to(torch.uint8)
    crop = torch.randint(0, 2 * pad + 1, (n, 2), generator=g, dtype=torch.int32)
    return flip, crop


def preprocess_index_map(h_in: int, w_in: int, out_h: int, out_w: int, pad: int, flip: int, top: int, left: int) -> np.ndarray:
    """ Integer source index (row, col) of every output pixel, (-1,-1) where the source is zero padding.
    RandomCrop(size, padding=pad, fill=0) geometry then hflip of the cropped image. """
    idx = np.full((out_h, out_w, 2), -1, dtype=np.int32)
    for i in range(out_h):
        for j in range(out_w):
            jj = out_w - 1 - j if flip else j  # out[i,j] = cropped[i, w-1-j]
            r, c = top + i - pad, left + jj - pad
            if 0 <= r < h_in and 0 <= c < w_in:
                idx[i, j] = (r, c)
    return idx


def preprocess_u8(img_u8_nhwc: torch.Tensor, mean: Sequence[float], std: Sequence[float], flip: torch.Tensor = None, crop_yx: torch.Tensor = None,
                  pad: int = 0, out_hw: Tuple[int, int] = None) -> torch.Tensor:
    """ uint8 N x H x W x C -> float32 N x C x h x w. `torchvision.transforms.ToTensor` then `Normalize`
    (conf/base/parameters.yml:197-210; meta/data/preprocess.py:44-57), same fp32 op order: (u8/255 - mean)/std. """
    n, h, w, c = img_u8_nhwc.shape
    oh, ow = out_hw if out_hw is not None else (h, w)
    src = F.pad(img_u8_nhwc.permute(0, 3, 1, 2), (pad, pad, pad, pad), value=0) if pad else img_u8_nhwc.permute(0, 3, 1, 2)
    out = torch.empty(n, c, oh, ow, dtype=torch.uint8)
    for k in range(n):
        top, left = (int(crop_yx[k, 0]), int(crop_yx[k, 1])) if crop_yx is not None else (pad, pad)
        patch = src[k, :, top:top + oh, left:left + ow]
        out[k] = patch.flip(-1) if (flip is not None and int(flip[k])) else patch
    x = out.to(torch.float32).div(255)
    mean_t = torch.tensor(mean, dtype=torch.float32).view(1, c, 1, 1)
    std_t = torch.tensor(std, dtype=torch.float32).view(1, c, 1, 1)
    return (x - mean_t) / std_t


# --------------------------------------------------------------------------------------------------------------------
# Training step (meta/ignite_training.py:233-255; classification/image.py:64-80)

def train_step(model: torch.nn.Module, x: torch.Tensor, y: torch.Tensor, optimizer: Optional[torch.optim.Optimizer] = None, loss_fn=None) -> Tuple[float, torch.Tensor]:
    loss_fn = loss_fn if loss_fn is not None else torch.nn.CrossEntropyLoss()
    model.train()
    y_pred = model(x)
    loss = loss_fn(y_pred, y)
    if optimizer is not None:
        optimizer.zero_grad()
    else:
        model.zero_grad()
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    return loss.item(), y_pred.detach()


def resnet_style_spec(num_classes_unused: int = 1000) -> Dict[str, Any]:
    """ The C4 ResNet-style hp dict is YAML (conf/base/resnet_style.yml); kept here only as a loader convenience. """
    from deepcv_b200.yaml_config import load_parameters, find_model_spec
    from pathlib import Path
    return find_model_spec(load_parameters(Path(__file__).resolve().parent.parent / 'conf' / 'base' / 'resnet_style.yml'), 'resnet_style_classifier')


# --------------------------------------------------------------------------------------------------------------------
# bf16-storage emulation: the oracle's arithmetic stays float32 (the reference CPU path), but every tensor the bf16 device path keeps in
# HBM as bfloat16 is rounded at the same point: convolution operands (weights), block outputs after the activation (y) and after the
# normalisations (z), pooling / link outputs, and on the way back the gradients w.r.t. pre-activations (dy) and block inputs (dx).
# This separates "bf16 storage changes the numbers" (inherent, identical here and on the device) from kernel defects.

class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd: bool, bwd: bool):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def _q(x, fwd=True, bwd=False):
    return _Round.apply(x, fwd, bwd)


def _few_channel_block(op, x_shape) -> bool:
    """ Mirror of `dcv_sc_conv_supported` (include/deepcv_b200.h): the blocks the device runs on its fused few-channel kernels. """
    if not isinstance(op, torch.nn.Conv2d) or len(x_shape) != 4:
        return False
    k, (r, s), (h, w) = op.out_channels, op.kernel_size, x_shape[2:]
    ok = r == s and r in (3, 5) and op.stride == (1, 1) and op.dilation == (1, 1) and op.padding == (r // 2, r // 2) and op.groups == 1
    ok = ok and w % 16 == 0 and w <= 64 and h <= 64 and (h * w) % 16 == 0 and (op.in_channels <= 4 or op.in_channels == 16) and k % 2 == 0 and (k <= 4 or k == 16)
    return ok and (r == 3 or (op.in_channels <= 4 and k <= 4))


def emulate_bf16_storage(model: torch.nn.Module) -> torch.nn.Module:
    """ Patches (in place) the forward of every block of an `OracleDeepcvModule` so that it rounds tensors where the device path stores bf16.
    A few-channel block directly followed by a non-overlapping average pooling (and not referenced by a later link) never stores its normalised output:
    the device pools the raw output in fp32 and applies the normalisation to the pooled value (pool(A*y + B) = A*pool(y) + B), one rounding later. The
    other convolution blocks do the same in front of a 2x2 / stride-2 pooling, and add the referenced tensor of a residual sum that directly follows
    inside their normalisation pass (no rounding of the normalised tensor in between). """
    for container in model.modules():
        subs = getattr(container, '_submodules', None)
        if isinstance(subs, dict):
            names, mods = list(subs.keys()), list(subs.values())
            referenced = {r for m in mods for r in (getattr(m, 'referenced_submodules', None) or [])}
            for name, m, nxt in zip(names, mods, mods[1:]):
                if isinstance(nxt, torch.nn.AvgPool2d) and name not in referenced:
                    ks, st = nxt.kernel_size, nxt.stride
                    ks, st = (ks, ks) if isinstance(ks, int) else tuple(ks), (st, st) if isinstance(st, int) else tuple(st)
                    m._emul_next_pool = ks[0] if (ks == st and ks[0] == ks[1]) else 0
                # any other block directly followed by a residual sum with ONE referenced tensor: the sum rides in the block's normalisation pass
                # (z = A*y + B + other, one rounding)
                if isinstance(nxt, Link) and nxt.reduction == 'sum' and len(nxt.referenced_submodules) == 1 and not nxt.ignore_input and name not in referenced:
                    m._emul_next_sum = True
    for m in model.modules():
        if isinstance(m, torch.nn.Sequential) and len(m) > 0 and any(isinstance(c, (torch.nn.modules.conv._ConvNd, torch.nn.Linear)) for c in m):
            def fwd(x, m=m):
                op = next(c for c in m if isinstance(c, (torch.nn.modules.conv._ConvNd, torch.nn.Linear)))
                rest = [c for c in m if c is not op and not isinstance(c, torch.nn.Dropout)]
                act = [c for c in rest if not isinstance(c, (torch.nn.modules.batchnorm._BatchNorm, torch.nn.GroupNorm))]
                norms = [c for c in rest if c not in act]
                x = _q(x, fwd=False, bwd=True)                         # dx is stored bf16
                if isinstance(op, torch.nn.Linear):
                    pre = F.linear(x, op.weight, op.bias)              # fp32 weights, fp32 logits
                    for a in act:
                        pre = a(pre)
                    return pre
                w = _q(op.weight, fwd=True, bwd=False)                 # bf16 convolution operand, fp32 master weight and gradient
                pre = F.conv2d(x, w, op.bias, op.stride, op.padding, op.dilation, op.groups)
                pre = _q(pre, fwd=False, bwd=True)                     # dy (gradient w.r.t. the pre-activation) is stored bf16
                for a in act:
                    pre = a(pre)
                y = _q(pre)                                            # y stored bf16; statistics are taken on the stored values
                if norms:
                    for nrm in norms:
                        y = nrm(y)
                    pool = getattr(m, '_emul_next_pool', 0)
                    few = _few_channel_block(op, x.shape)
                    # few-channel blocks fuse any non-overlapping pooling; the other blocks the 2x2 / stride-2 one (and a residual sum)
                    fused_pool = pool and (few or pool == 2) and y.shape[2] % pool == 0 and y.shape[3] % pool == 0
                    fused_sum = getattr(m, '_emul_next_sum', False) and not few and isinstance(op, torch.nn.Conv2d)
                    if not (fused_pool or fused_sum):
                        y = _q(y)                                      # z stored bf16 (or rounded to bf16 when the next few-channel block stages it)
                return y
            m.forward = fwd
        elif isinstance(m, (torch.nn.AvgPool2d,)):
            orig = m.forward
            m.forward = (lambda x, orig=orig: _q(orig(_q(x, fwd=False, bwd=True))))
        elif isinstance(m, Link):
            orig_l = m.forward
            m.forward = (lambda x, referenced_submodules_out, orig_l=orig_l: _q(orig_l(_q(x, fwd=False, bwd=True), OrderedDict((k, _q(v, fwd=False, bwd=True)) for k, v in referenced_submodules_out.items()))))
    return model
