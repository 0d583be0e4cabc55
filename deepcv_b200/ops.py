""" Torch-facing operators of the hot path: `torch.autograd.Function`s whose forward and backward are calls into the C ABI
(`include/deepcv_b200.h`, bound in `deepcv_b200/_lib.py`). PyTorch is used here for device memory (caching allocator),
the current stream and the autograd graph only; every arithmetic pass is one of the library's CUDA kernels.

Conventions
  * Image tensors are logically `N x C x H x W` (what the reference's modules see) and physically NHWC
    (`memory_format=torch.channels_last`); `empty_nhwc` / `as_nhwc` create / obtain that layout.
  * Activation dtype = dtype of the incoming tensor (float32: parity mode, bfloat16: throughput mode). Parameters, statistics,
    parameter gradients and losses are float32.
  * No CPU implementation exists: a non-CUDA tensor raises `RuntimeError` (tensors on the `meta` device are handled by the
    modules in `deepcv_b200.meta.nn`, for shape inference only).
"""
import contextlib
import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import (ACT_LEAKY_RELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ALGO_AUTO, ALGO_DIRECT, DCV_BF16, DCV_F32, ConvShape, LINK_MAX_SOURCES, LinkSource, NormParams, ScNorm, check, lib)

__all__ = ['empty_nhwc', 'is_nhwc', 'as_nhwc', 'activation_code', 'conv_block', 'NormConfig', 'avg_pool2d', 'link_reduce', 'link_concat_rescaled', 'bilinear_resize', 'flatten_nchw',
           'linear_act', 'cross_entropy', 'preprocess_u8', 'fork', 'launch_count', 'AccumulatorArena', 'StepContext', 'PendingAffine', 'sc_conv_block',
           'sc_conv_supported', 'sc_affine_pool', 'materialize', 'DropoutState', 'dropout', 'activation', 'pre_norm_act', 'PendingFlatten', 'unit_grad', 'PendingNorm', 'apply_pending']

_DTYPES = {torch.float32: DCV_F32, torch.bfloat16: DCV_BF16}
_DEBUG_CAPTURE = None   # tests may set this to a list to capture backward intermediates of conv blocks


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f'deepcv_b200: unsupported dtype {t.dtype} (float32 and bfloat16 only)') from None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(f'deepcv_b200: got a tensor on "{t.device}"; this path runs on CUDA (sm_100a) only and has no CPU fallback')


def launch_count() -> int:
    """ Number of kernels launched through the library since it was loaded. """
    return int(lib.dcv_launch_count())


def empty_nhwc(n: int, c: int, h: int, w: int, dtype: torch.dtype, device) -> torch.Tensor:
    return torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def is_nhwc(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous()


class _ToChannelsLast(torch.autograd.Function):
    """ contiguous NCHW (any supported dtype) -> NHWC `dtype`, through the tiled transpose kernel. """

    @staticmethod
    def forward(ctx, x: torch.Tensor, dtype: torch.dtype):
        n, c, h, w = x.shape
        out = empty_nhwc(n, c, h, w, dtype, x.device)
        check(lib.dcv_nchw_to_nhwc(_ptr(x), _dt(x), _ptr(out), _dt(out), n, c, h, w, _stream()), 'nchw_to_nhwc')
        ctx.src_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        g = as_nhwc(g)
        n, c, h, w = g.shape
        dx = torch.empty((n, c, h, w), dtype=ctx.src_dtype, device=g.device)
        check(lib.dcv_nhwc_to_nchw(_ptr(g), _dt(g), _ptr(dx), _dt(dx), n, c, h, w, _stream()), 'nhwc_to_nchw')
        return dx, None


class _Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, dtype: torch.dtype):
        out = torch.empty_like(x, dtype=dtype)  # preserves the (dense) strides
        check(lib.dcv_cast(_ptr(x), _dt(x), _ptr(out), _dt(out), x.numel(), _stream()), 'cast')
        ctx.src_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        return _cast_raw(_dense_like(g), ctx.src_dtype), None


def _dense_like(t: torch.Tensor) -> torch.Tensor:
    """ A tensor whose memory is dense in either NHWC or plain contiguous order (autograd may hand us expanded gradients). """
    if t.is_contiguous() or (t.dim() == 4 and is_nhwc(t)):
        return t
    return t.contiguous()


def _cast_raw(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if x.dtype == dtype:
        return x
    out = torch.empty_like(x, dtype=dtype)
    check(lib.dcv_cast(_ptr(x), _dt(x), _ptr(out), _dt(out), x.numel(), _stream()), 'cast')
    return out


def as_nhwc(x: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """ `x` as a physically-NHWC tensor of `dtype` (default: unchanged), differentiable. """
    _require_cuda(x)
    if x.dim() != 4:
        raise ValueError(f'deepcv_b200: expected an N x C x H x W tensor, got shape {tuple(x.shape)}')
    dtype = x.dtype if dtype is None else dtype
    if is_nhwc(x):
        return x if x.dtype == dtype else _Cast.apply(x, dtype)
    if not x.is_contiguous():
        x = x.contiguous()
    return _ToChannelsLast.apply(x, dtype)


def activation_code(act_fn) -> Tuple[int, float]:
    """ Activation module / type / None -> (DCV_ACT_*, negative slope). Only what the kernels fuse is accepted. """
    if act_fn is None or act_fn is torch.nn.Identity or isinstance(act_fn, torch.nn.Identity):
        return ACT_NONE, 0.
    inst = act_fn() if isinstance(act_fn, type) else act_fn
    if isinstance(inst, torch.nn.ReLU):
        return ACT_RELU, 0.
    if isinstance(inst, torch.nn.LeakyReLU):
        return ACT_LEAKY_RELU, float(inst.negative_slope)
    if isinstance(inst, torch.nn.Sigmoid):
        return ACT_SIGMOID, 0.
    raise NotImplementedError(f'deepcv_b200: activation "{type(inst).__name__}" is not fused by the sm_100a kernels (ReLU, LeakyReLU, Sigmoid, Identity/None are); '
                              'there is no PyTorch fallback on this path')


# ------------------------------------------------------------------------------------------------------------------------------
# Convolution block: conv (+bias) -> activation -> [BatchNorm] -> [GroupNorm / InstanceNorm]   (reference: meta/nn.py:553)

class NormConfig:
    """ What `dcv_norm_params` needs besides tensors: which normalisations follow the activation and their scalars. """
    __slots__ = ('use_bn', 'bn_eps', 'bn_momentum', 'use_gn', 'gn_groups', 'gn_eps')

    def __init__(self, use_bn=False, bn_eps=1e-5, bn_momentum=0.1, use_gn=False, gn_groups=1, gn_eps=1e-5):
        self.use_bn, self.bn_eps, self.bn_momentum = bool(use_bn), float(bn_eps), (-1. if bn_momentum is None else float(bn_momentum))
        self.use_gn, self.gn_groups, self.gn_eps = bool(use_gn), int(gn_groups), float(gn_eps)

    @property
    def any(self) -> bool:
        return self.use_bn or self.use_gn


class AccumulatorArena:
    """ One zeroed buffer per training step for every accumulator the kernels fill with atomics (per-(n,c) statistics, backward sums, bias / weight
    gradients that are not written straight into the flat gradient buffer): `begin_step` zeroes it (and any extra tensors, e.g. the flat gradient
    buffer) with ONE memset each; while it is active the accumulating entry points are called with `acc_prezeroed = 1` — a CIFAR step otherwise
    carries ~22 memset nodes of 2-3 us. Used by `GraphedTrainStep` around the captured step; outside `begin_step` .. `end_step` nothing changes.
    `measure()` .. `end_measure()` runs a step in counting mode to size the buffer. An allocation that does not fit raises (never a silent fallback).
    One instance per captured step: the graph replays keep writing into `buf`. Reached through the `StepContext` of the layers, never a global. """

    def __init__(self):
        self.buf, self.off, self.active, self.counting, self.need = None, 0, False, False, 0

    def measure(self):
        self.counting, self.need = True, 0

    def end_measure(self, device) -> int:
        self.counting = False
        if self.buf is None or self.buf.numel() < self.need or self.buf.device != device:
            self.buf = torch.empty((max(self.need, 256),), dtype=torch.uint8, device=device)
        return self.need

    def begin_step(self, extra_zero=()):
        if self.buf is None:
            raise RuntimeError('deepcv_b200: AccumulatorArena.begin_step before measure() / end_measure()')
        st = _stream()
        check(lib.dcv_fill_zero(_ptr(self.buf), self.buf.numel(), st), 'fill_zero(arena)')
        for t in extra_zero:
            check(lib.dcv_fill_zero(_ptr(t), t.numel() * t.element_size(), st), 'fill_zero(gradients)')
        self.off, self.active = 0, True

    def end_step(self):
        self.active = False

    def alloc(self, shape, device) -> torch.Tensor:
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = (numel * 4 + 255) // 256 * 256
        if self.active:
            if self.off + nbytes > self.buf.numel():
                raise RuntimeError(f'deepcv_b200: accumulator arena exhausted ({self.off} + {nbytes} > {self.buf.numel()} bytes): the step changed shape since it was measured')
            out = self.buf[self.off:self.off + numel * 4].view(torch.float32).view(*shape)
            self.off += nbytes
            return out
        if self.counting:
            self.need += nbytes
        return torch.empty(tuple(shape), dtype=torch.float32, device=device)


class StepContext:
    """ Per-step resources of ONE model's training step, handed to the fused layers by whoever drives the step (`GraphedTrainStep` sets
    `FusedLayer._step_ctx` on the layers of its model for the duration of warm-up + capture): the accumulator arena and the low-precision
    shadow of the flat parameter buffer. It travels as an argument through the autograd functions (forward thread and autograd's worker thread
    see the same object), so two models, streams or capturing threads never share state — there is no module-level "current" anything. """

    def __init__(self, arena: Optional[AccumulatorArena] = None):
        self.arena = arena
        self.shadows: List[Tuple[torch.Tensor, torch.Tensor]] = []   # (flat fp32 parameter buffer, same-length buffer of the operand dtype)
        self.transposed = None   # see `plan_transposed_weights`
        self.side_stream: Optional[torch.cuda.Stream] = None   # weight-gradient kernels run here, concurrently with the rest of backward (`_ConvBlock.backward`)
        self.keep_alive: list = []

    def join_side(self) -> None:
        """ End of backward: the compute stream waits for the weight-gradient kernels on the side stream; the tensors they used may be released. """
        if self.side_stream is not None:
            torch.cuda.current_stream().wait_stream(self.side_stream)
        self.keep_alive.clear()

    @property
    def prezeroed(self) -> int:
        """ 1 iff accumulators come from the arena that the step zeroes up front (and so does the flat gradient buffer). """
        return int(self.arena is not None and self.arena.active)

    def acc_empty(self, shape, device) -> torch.Tensor:
        if self.arena is not None:
            return self.arena.alloc(shape, device)
        return torch.empty(tuple(int(d) for d in shape), dtype=torch.float32, device=device)

    def refresh_shadows(self) -> None:
        """ ONE cast kernel per step instead of one per convolution. """
        for src, dst in self.shadows:
            check(lib.dcv_cast(_ptr(src), _dt(src), _ptr(dst), _dt(dst), src.numel(), _stream()), 'cast(flat parameters)')
        t = self.transposed
        if t is not None:
            check(lib.dcv_pack_conv_weights_batched(_ptr(t['flat']), _ptr(t['buf']), _dt(t['buf']), _ptr(t['table']), t['n'], t['units'], _stream()), 'pack_conv_weights_batched')

    def plan_transposed_weights(self, flat_params: torch.Tensor, weights: Sequence[torch.Tensor], dtype: torch.dtype) -> None:
        """ The data-gradient operands ([C][R-1-r][S-1-s][K], operand dtype) of all the given convolution weights — [K][R][S][C] slices of `flat_params` —
        are produced by ONE kernel per step (`refresh_shadows`) into one buffer, instead of one `dcv_pack_conv_weight` launch per layer in backward. """
        from ._lib import PackEntry
        entries, views, dst_off, unit0 = [], {}, 0, 0
        for w in weights:
            k, c, r, s_ = w.shape
            off = (w.data_ptr() - flat_params.data_ptr()) // 4
            if not (w.dtype == torch.float32 and w.permute(0, 2, 3, 1).is_contiguous() and 0 <= off and off + w.numel() <= flat_params.numel()):
                continue
            entries.append(PackEntry(off, dst_off, unit0, k, r, s_, c))
            views[w.data_ptr()] = (dst_off, (c, r, s_, k))
            dst_off += (w.numel() + 7) // 8 * 8      # 16-byte aligned slices (TMA)
            unit0 += r * s_ * ((c + 7) // 8) * ((k + 31) // 32)
        if not entries or len(entries) > 64:
            return
        table = (PackEntry * len(entries))(*entries)
        dev_table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(flat_params.device)
        buf = torch.empty(dst_off, dtype=dtype, device=flat_params.device)
        self.transposed = dict(flat=flat_params, table=dev_table, n=len(entries), units=unit0, buf=buf, views=views)

    def transposed_view(self, weight: torch.Tensor, dtype: torch.dtype) -> Optional[torch.Tensor]:
        t = self.transposed
        if t is None or t['buf'].dtype != dtype:
            return None
        hit = t['views'].get(weight.data_ptr())
        if hit is None:
            return None
        off, shape = hit
        n = shape[0] * shape[1] * shape[2] * shape[3]
        return t['buf'][off:off + n].view(*shape)

    def shadow_view(self, weight: torch.Tensor, dtype: torch.dtype) -> Optional[torch.Tensor]:
        for src, dst in self.shadows:
            if dst.dtype != dtype or weight.dtype != src.dtype or weight.device != src.device:
                continue
            off = weight.data_ptr() - src.data_ptr()
            if 0 <= off and off + weight.numel() * 4 <= src.numel() * 4 and off % 4 == 0:
                k, c, r, s_ = weight.shape
                return dst[off // 4: off // 4 + weight.numel()].view(k, r, s_, c).permute(0, 3, 1, 2)
        return None


_NO_CTX = StepContext()   # eager calls outside a driven step: plain allocations, entry points zero their accumulators, per-layer weight casts


def _pz(sctx: Optional['StepContext']) -> int:
    return 0 if sctx is None else sctx.prezeroed


def _acc_empty(shape, device, sctx: Optional['StepContext'] = None) -> torch.Tensor:
    """ fp32 accumulator buffer: from the step's zeroed arena when one is active, else plain (the filling entry point zeroes it). """
    return (sctx or _NO_CTX).acc_empty(shape, device)


# DCV_FUSED_STATS=1: BatchNorm-only blocks take their batch statistics from the tcgen05 epilogue (DCV_STATS_CHANNEL_TOTALS) instead of a statistics pass
# over y. Built, bit-exact (tests/test_gpu_exact.py) and OFF by default: measured on the ImageNet-shaped step (B200, round 2) it removes 18 launches /
# 0.45 ms of statistics kernels but lengthens the convolution epilogues — which are on these kernels' critical path — by more: 7.05 -> 7.22 ms
# (64-channel halo kernel 77 -> 90 us with per-thread partial sums, 128 / 256-channel kernels 60 -> 90 us with a per-tile butterfly, stem 402 -> 594 us).
_FUSE_STATS = os.environ.get('DCV_FUSED_STATS') == '1'
_CHANNEL_TOTALS = os.environ.get('DCV_NO_CHANNEL_TOTALS') is None   # A/B switch: per-(image, channel) sums even for BatchNorm-only blocks
_SIDE_WGRAD = os.environ.get('DCV_SIDE_WGRAD') == '1'   # opt-in: weight-gradient kernels on a second stream (measured: 6.377 -> 6.323 ms on the ImageNet-shaped step — they compete with the normalisation passes for HBM — not worth a second stream inside the captured step by default)
_FOLD_FWD_FINALIZE = os.environ.get('DCV_NO_FOLD_FWD_FINALIZE') is None   # tuning aid: DCV_NO_FOLD_FWD_FINALIZE=1 keeps the stand-alone forward finalize launch of BatchNorm-only blocks
_FOLD_BWD_FINALIZE = os.environ.get('DCV_NO_FOLD_BWD_FINALIZE') is None   # tuning aid: DCV_NO_FOLD_BWD_FINALIZE=1 keeps the stand-alone backward finalize launch of BatchNorm-only blocks
_ONE_IMAGE_BN = os.environ.get('DCV_NO_ONE_IMAGE_BN') is None   # tuning aid: DCV_NO_ONE_IMAGE_BN=1 keeps per-image coefficient tables for BatchNorm-only blocks
_POOLED_BWD = os.environ.get('DCV_NO_POOLED_BWD') is None   # tuning aid: DCV_NO_POOLED_BWD=1 materialises the full-resolution gradient behind a fused normalise + pool
_LAZY_APPLY = os.environ.get('DCV_NO_LAZY_APPLY') is None   # tuning aid: DCV_NO_LAZY_APPLY=1 always runs the stand-alone normalisation apply pass
_USE_PAIRS = os.environ.get('DCV_NO_PAIRS') is None     # tuning aid: DCV_NO_PAIRS=1 sends stride-2 few-channel layers to the gather kernels instead of the pixel-pair ones
_USE_GATHER = os.environ.get('DCV_NO_GATHER') is None   # tuning aid: DCV_NO_GATHER=1 forces the explicit im2col route for the stem


def _conv_shape(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation) -> ConvShape:
    n, c, h, w = x.shape
    k, cw, r, s = weight.shape
    if cw != c:
        raise RuntimeError(f'deepcv_b200: convolution expects {cw} input channels, got {c} (grouped convolutions are not on this path)')
    p = (h + 2 * padding[0] - dilation[0] * (r - 1) - 1) // stride[0] + 1
    q = (w + 2 * padding[1] - dilation[1] * (s - 1) - 1) // stride[1] + 1
    if p <= 0 or q <= 0:
        raise RuntimeError(f'deepcv_b200: convolution output would be empty for input {tuple(x.shape)} and kernel {r}x{s}')
    return ConvShape(n, h, w, c, k, r, s, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], p, q)


def _weight_operand(weight: torch.Tensor, dtype: torch.dtype, sctx: Optional[StepContext] = None) -> torch.Tensor:
    """ fp32 OIHW parameter -> [K][R][S][C] tensor of the activation dtype (zero-copy when it already is one, or a view of the step's
    low-precision shadow of the flat parameter buffer). """
    k, c, r, s = weight.shape
    if weight.permute(0, 2, 3, 1).is_contiguous():
        if weight.dtype == dtype:
            return weight
        shadow = sctx.shadow_view(weight, dtype) if (sctx is not None and sctx.shadows) else None
        if shadow is not None:
            return shadow
        out = torch.empty((k, r, s, c), dtype=dtype, device=weight.device).permute(0, 3, 1, 2)
        check(lib.dcv_cast(_ptr(weight), _dt(weight), _ptr(out), _dt(out), weight.numel(), _stream()), 'cast(weight)')
        return out
    w = weight if weight.is_contiguous() else weight.contiguous()
    out = torch.empty((k, r, s, c), dtype=dtype, device=weight.device).permute(0, 3, 1, 2)
    check(lib.dcv_nchw_to_nhwc(_ptr(w), _dt(w), _ptr(out), _dt(out), k, c, r, s, _stream()), 'nchw_to_nhwc(weight)')
    return out


def _norm_params(cfg: NormConfig, n: int, c: int, hw: int, training: bool, bn_w, bn_b, rm, rv, nbt, gn_w, gn_b) -> NormParams:
    bn_training = bool(training or rm is None or rv is None)  # track_running_stats=False => always batch statistics
    return NormParams(n, c, hw, int(cfg.use_bn), int(bn_training), cfg.bn_eps, cfg.bn_momentum,
                      _ptr(bn_w), _ptr(bn_b), _ptr(rm), _ptr(rv), _ptr(nbt),
                      int(cfg.use_gn), cfg.gn_groups, cfg.gn_eps, _ptr(gn_w), _ptr(gn_b))


def _bn_apply_fold(fold, y, other, pool: bool, out, rows: int, w: int, c: int, dt: int, st) -> None:
    stats, bn_w, bn_b, rm, rv, nbt, eps, momentum, bn_training, saved = fold
    check(lib.dcv_bn_apply_fold_fwd(_ptr(y), _ptr(other), int(pool), _ptr(out), _ptr(stats), _ptr(bn_w), _ptr(bn_b), _ptr(rm), _ptr(rv), _ptr(nbt), eps, momentum, bn_training,
                                    _ptr(saved), rows, w, c, dt, st), 'bn_apply_fold_fwd')


class _ConvBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, bn_w, bn_b, gn_w, gn_b, rm, rv, nbt, stride, padding, dilation, act, slope, cfg: NormConfig, training: bool, algo: int, grad_out, sctx, notify=True, defer_apply=False, link=None):
        _require_cuda(x, weight)
        ctx.notify = notify
        ctx.link = link   # side channel from the consumer of a pending normalisation (see `_ApplyNorm.backward`)
        ctx.set_materialize_grads(False)   # the coefficient table returned next to a deferred output has no gradient: no zero-fill kernel for it in backward
        shape = _conv_shape(x, weight, stride, padding, dilation)
        n, k, p, q = shape.n, shape.k, shape.p, shape.q
        dev, st = x.device, _stream()
        dt = _dt(x)
        pz = _pz(sctx)
        w_op = _weight_operand(weight, x.dtype, sctx)
        y = empty_nhwc(n, k, p, q, x.dtype, dev)
        stats = _acc_empty((n, k, 2), dev, sctx) if cfg.any else None
        # BatchNorm alone needs per-channel totals only (DCV_STATS_CHANNEL_TOTALS): the batch is reduced as one image; opt-in: in the tcgen05 epilogue
        totals = (2 | (4 if _FUSE_STATS else 0)) if (cfg.use_bn and not cfg.use_gn and _CHANNEL_TOTALS) else 0
        # Convolutions the implicit-GEMM tensor-core kernel cannot address directly (few input channels / strides: the 7x7 stride-2 stem) go
        # through an explicit im2col: conv(x, w) == 1x1 conv of col[n][p][q][kpad] with the weights zero-padded to [K][kpad].
        rsc = shape.r * shape.s * shape.c
        gemm_shape = None
        if (x.dtype == torch.bfloat16 and algo != ALGO_DIRECT and k % 64 == 0 and rsc <= 4096 and n * p * q >= 128
                and not lib.dcv_conv2d_tc_supported(ctypes.byref(shape), dt, 0)):
            kpad = (rsc + 63) // 64 * 64
            gemm_shape = ConvShape(n, p, q, kpad, k, 1, 1, 1, 1, 0, 0, 1, 1, p, q)
            if not lib.dcv_conv2d_tc_supported(ctypes.byref(gemm_shape), dt, 0):
                gemm_shape = None
        gathered = 0
        if gemm_shape is not None:
            # gather kernels: producer warps build the im2col tile in shared memory, col[n][p][q][kpad] (1.2 GB for the stem at batch 256) is never
            # materialised. Their K order pads every filter row to a multiple of 8 elements (see dcv_gather_pack_weight).
            sc = shape.s * shape.c
            kpad_g = (shape.r * ((sc + 7) // 8 * 8) + 63) // 64 * 64
            if _USE_GATHER and _USE_PAIRS and lib.dcv_conv2d_pairs_supported(ctypes.byref(shape), _ptr(x), dt):
                gathered = 2     # stride 2, at most four input channels (the stem): tensor-core tiles straight over pair-transposed input rows
            elif _USE_GATHER and lib.dcv_conv2d_gather_supported(ctypes.byref(shape), _ptr(x), kpad_g, dt):
                gathered = 1
            if gathered == 2:
                w_col = torch.empty((k, 256), dtype=x.dtype, device=dev)
                check(lib.dcv_pairs_pack_weight(_ptr(w_op), _ptr(w_col), ctypes.byref(shape), dt, st), 'pairs_pack_weight')
                check(lib.dcv_conv2d_fwd_pairs(ctypes.byref(shape), _ptr(x), _ptr(w_col), _ptr(bias), _ptr(y), _ptr(stats), act, slope, pz | totals, st), 'conv2d_fwd_pairs')
            elif gathered:
                w_col = torch.empty((k, kpad_g), dtype=x.dtype, device=dev)
                check(lib.dcv_gather_pack_weight(_ptr(w_op), _ptr(w_col), k, shape.r, sc, kpad_g, dt, st), 'gather_pack_weight')
                check(lib.dcv_conv2d_fwd_gather(ctypes.byref(shape), _ptr(x), _ptr(w_col), kpad_g, _ptr(bias), _ptr(y), _ptr(stats), act, slope, pz | totals, st), 'conv2d_fwd_gather')
            else:
                w_col = torch.empty((k, kpad), dtype=x.dtype, device=dev)
                check(lib.dcv_fill_zero(_ptr(w_col), w_col.numel() * w_col.element_size(), st), 'fill_zero')
                check(lib.dcv_copy_channels_in(_ptr(w_op), _ptr(w_col), k, rsc, kpad, 0, dt, st), 'copy_channels_in(weight)')
                col = torch.empty((n, p, q, kpad), dtype=x.dtype, device=dev)
                check(lib.dcv_im2col(ctypes.byref(shape), _ptr(x), _ptr(col), kpad, dt, st), 'im2col')
                check(lib.dcv_conv2d_fwd(ctypes.byref(gemm_shape), _ptr(col), _ptr(w_col), _ptr(bias), _ptr(y), _ptr(stats), act, slope, dt, algo, pz | totals, st), 'conv2d_fwd(im2col)')
                x = col   # what the weight gradient needs; the data gradient only needs dy and the weights
        else:
            check(lib.dcv_conv2d_fwd(ctypes.byref(shape), _ptr(x), _ptr(w_op), _ptr(bias), _ptr(y), _ptr(stats), act, slope, dt, algo, pz | totals, st), 'conv2d_fwd')
        saved = None
        out = y
        # BatchNorm-only block whose statistics came out as channel totals (tensor-core paths): to the finalize / apply kernels the batch IS one image of
        # n*p*q pixels (same arithmetic: BatchNorm reduces over (N, H, W)) — coefficient tables of k entries instead of n*k, finalize kernels that do not
        # walk n rows of zeros
        one_image = bool(totals & 2) and n > 1 and n * p * q < (1 << 31) and _ONE_IMAGE_BN and \
            bool(gathered or gemm_shape is not None or (algo != ALGO_DIRECT and lib.dcv_conv2d_tc_supported(ctypes.byref(shape), dt, 0)))
        ne, hwe = (1, n * p * q) if one_image else (n, p * q)
        fold = None
        if cfg.any:
            groups = cfg.gn_groups if cfg.use_gn else 1
            saved = torch.empty((int(lib.dcv_norm_saved_floats(ne, k, groups)),), dtype=torch.float32, device=dev)
            # BatchNorm-only block handed over as one image: no finalize launch — the kernel that applies the normalisation (here, or the consumer of the
            # pending normalisation) computes the coefficients from the channel totals in its prologue and writes `saved` / the running statistics
            if one_image and cfg.use_bn and not cfg.use_gn and cfg.bn_momentum >= 0. and _FOLD_FWD_FINALIZE and k % (16 // y.element_size()) == 0:
                bn_training = bool(training or rm is None or rv is None)
                fold = (stats, bn_w, bn_b, rm, rv, nbt, float(cfg.bn_eps), float(cfg.bn_momentum), int(bn_training), saved)
                ab = saved   # what travels with a pending normalisation (`PendingNorm.ab`): the consumer finds the fold arguments in the link
                if link is not None:
                    link['fold'] = fold
                if not defer_apply:
                    out = empty_nhwc(n, k, p, q, x.dtype, dev)
                    _bn_apply_fold(fold, y, None, False, out, n * p, q, k, dt, st)
            else:
                ab = torch.empty((ne, k, 2), dtype=torch.float32, device=dev)
                prm = _norm_params(cfg, ne, k, hwe, training, bn_w, bn_b, rm, rv, nbt, gn_w, gn_b)
                check(lib.dcv_norm_fwd_finalize(ctypes.byref(prm), _ptr(stats), _ptr(ab), _ptr(saved), st), 'norm_fwd_finalize')
                if not defer_apply:
                    out = empty_nhwc(n, k, p, q, x.dtype, dev)
                    check(lib.dcv_norm_apply_fwd(_ptr(y), _ptr(ab), _ptr(out), ne, hwe, k, dt, st), 'norm_apply_fwd')
        ctx.save_for_backward(x, w_op, y, stats, saved, bn_w, bn_b, gn_w, gn_b, rm, rv)
        # the fp32 master weight, when it already is [K][R][S][C] in memory: the data-gradient operand is packed straight from it in backward
        ctx.w_master = weight.detach() if (weight.dtype == torch.float32 and weight.permute(0, 2, 3, 1).is_contiguous()) else None
        ctx.cfg = (shape, act, slope, cfg, training, algo, bias is not None, grad_out, tuple(weight.shape), gemm_shape, gathered, sctx)
        ctx.one_image = one_image
        if defer_apply and cfg.any:
            # the RAW output goes on with its coefficients (`PendingNorm`): the consumer applies z = A*y + B inside its own pass (`_ApplyNorm*`) and hands
            # back dz — the gradient w.r.t. the normalised output, exactly what this backward expects — as the "gradient" of y
            ctx.mark_non_differentiable(ab)
            return y, ab
        return out

    @staticmethod
    def backward(ctx, dz, _dab=None):
        x, w_op, y, stats, saved, bn_w, bn_b, gn_w, gn_b, rm, rv = ctx.saved_tensors
        shape, act, slope, cfg, training, algo, has_bias, grad_out, wshape, gemm_shape, gathered, sctx = ctx.cfg
        pz = _pz(sctx)
        n, k, p, q = shape.n, shape.k, shape.p, shape.q
        dev, st, dt = y.device, _stream(), _dt(y)
        # a 2x2 pooling consumed the pending normalisation: its gradient arrives at POOLED resolution through the side channel (the `dz` argument is a
        # placeholder of the right shape) and the two normalisation passes below read it in place
        dzp = ctx.link.pop('pooled_dz', None) if ctx.link is not None else None
        if dzp is None:
            if dz is None:   # (gradients are not materialised: an output nobody differentiates through)
                dz = torch.zeros_like(y)
            dz = as_nhwc(dz.detach(), y.dtype)
        f32 = dict(dtype=torch.float32, device=dev)
        # gradients go straight into the caller's bucket slices on the FIRST backward after `zero_grad`; a second backward through the same layer before the
        # next `zero_grad` (weights shared between two calls, gradient accumulation) returns tensors instead, which autograd accumulates into `.grad`
        targets = grad_out if (grad_out and not grad_out.get('_written', False)) else {}
        pqr = d_bn_w = d_bn_b = d_gn_w = d_gn_b = None
        if cfg.any:
            s_nc = _acc_empty((n, k, 2), dev, sctx)
            totals = 2 if (cfg.use_bn and not cfg.use_gn and _CHANNEL_TOTALS) else 0
            if dzp is not None:
                check(lib.dcv_norm_bwd_reduce_pooled(_ptr(dzp), _ptr(y), _ptr(s_nc), n, p, q, k, dt, pz | totals, st), 'norm_bwd_reduce_pooled')
            else:
                check(lib.dcv_norm_bwd_reduce(_ptr(dz), _ptr(y), _ptr(s_nc), n, p * q, k, dt, pz | totals, st), 'norm_bwd_reduce')
            ne, hwe = (1, n * p * q) if ctx.one_image else (n, p * q)
            pqr = torch.empty((ne, k, 3), **f32)
            if cfg.use_bn and bn_w is not None:
                d_bn_w, d_bn_b = targets.get('bn_w', None), targets.get('bn_b', None)
                d_bn_w = torch.empty((k,), **f32) if d_bn_w is None else d_bn_w
                d_bn_b = torch.empty((k,), **f32) if d_bn_b is None else d_bn_b
            if cfg.use_gn and gn_w is not None:
                d_gn_w, d_gn_b = targets.get('gn_w', None), targets.get('gn_b', None)
                d_gn_w = torch.empty((k,), **f32) if d_gn_w is None else d_gn_w
                d_gn_b = torch.empty((k,), **f32) if d_gn_b is None else d_gn_b
            # BatchNorm-only block handed over as one image: P, Q, R and the BatchNorm parameter gradients come out of the apply kernel's prologue (no finalize launch)
            fold = ctx.one_image and cfg.use_bn and not cfg.use_gn and _FOLD_BWD_FINALIZE and k % (16 // y.element_size()) == 0
            if not fold:
                prm = _norm_params(cfg, ne, k, hwe, training, bn_w, bn_b, rm, rv, None, gn_w, gn_b)
                check(lib.dcv_norm_bwd_finalize(ctypes.byref(prm), _ptr(stats), _ptr(s_nc), _ptr(saved), _ptr(pqr),
                                                _ptr(d_bn_w), _ptr(d_bn_b), _ptr(d_gn_w), _ptr(d_gn_b), st), 'norm_bwd_finalize')
        need_dy_pass = cfg.any or act != ACT_NONE or has_bias
        dy = dz
        dbias = None
        if need_dy_pass:
            dy = empty_nhwc(n, k, p, q, y.dtype, dev)
            if has_bias:
                dbias = targets.get('bias', None)
                dbias = _acc_empty((k,), dev, sctx) if dbias is None else dbias
            one = cfg.any and ctx.one_image   # the coefficient table has one row: the batch is one image of n*p rows
            if cfg.any and fold:
                bn_training = bool(training or rm is None or rv is None)
                check(lib.dcv_act_bn_bwd_apply_fold(_ptr(dzp if dzp is not None else dz), int(dzp is not None), _ptr(y), _ptr(s_nc), _ptr(saved), int(bn_training),
                                                    _ptr(d_bn_w), _ptr(d_bn_b), _ptr(dy), _ptr(dbias), act, slope, n * p, q, k, dt, pz, st), 'act_bn_bwd_apply_fold')
            elif dzp is not None:
                check(lib.dcv_act_norm_bwd_apply_pooled(_ptr(dzp), _ptr(y), _ptr(pqr), _ptr(dy), _ptr(dbias), act, slope, 1 if one else n, n * p if one else p, q, k, dt, pz, st),
                      'act_norm_bwd_apply_pooled')
            else:
                check(lib.dcv_act_norm_bwd_apply(_ptr(dz), _ptr(y), _ptr(pqr), _ptr(dy), _ptr(dbias), act, slope, 1 if one else n, n * p * q if one else p * q, k, dt, pz, st),
                      'act_norm_bwd_apply')
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty_nhwc(n, shape.c, shape.h, shape.w, y.dtype, dev)
            wt = None
            if algo != ALGO_DIRECT and lib.dcv_conv2d_tc_supported(ctypes.byref(shape), dt, 1):
                # operand of the data-gradient convolution: [C][R-1-r][S-1-s][K] in the activation dtype — from the step's batched pack when there is one
                wt = sctx.transposed_view(ctx.w_master, y.dtype) if (sctx is not None and ctx.w_master is not None) else None
            if wt is None and algo != ALGO_DIRECT and lib.dcv_conv2d_tc_supported(ctypes.byref(shape), dt, 1):
                wt = torch.empty((shape.c, shape.r, shape.s, shape.k), dtype=y.dtype, device=dev)
                # NB: packing from the fp32 master rounds to bf16 once, exactly like the forward operand (a bf16 -> fp32 -> bf16 round trip is the identity)
                w32 = ctx.w_master if ctx.w_master is not None else (_cast_raw(w_op, torch.float32) if w_op.dtype != torch.float32 else w_op)
                check(lib.dcv_pack_conv_weight(_ptr(w32), _ptr(wt), dt, shape.k, shape.r, shape.s, shape.c, 1, st), 'pack_conv_weight')
            check(lib.dcv_conv2d_dgrad(ctypes.byref(shape), _ptr(dy), _ptr(w_op), _ptr(wt), _ptr(dx), dt, algo, st), 'conv2d_dgrad')

        # The weight gradient goes LAST and, inside a step that provides a side stream (`StepContext.side_stream`), on that stream: it only feeds the optimizer,
        # while the data gradient is on the critical path to the previous layer — whose normalisation backward (HBM-bound reduce / finalize / apply passes)
        # then runs concurrently with this tensor-bound kernel. The step joins the side stream before it touches the gradients (`StepContext.join_side`).
        dw = side = None
        if ctx.needs_input_grad[1]:
            dw = targets.get('weight', None)
            # only a gradient written in place into the step's flat buffer (nothing handed to autograd), by the node that closes the layer's backward
            side = sctx.side_stream if (sctx is not None and dw is not None and ctx.notify and _SIDE_WGRAD and algo != ALGO_DIRECT) else None
            if dw is None:
                dw = _acc_empty((k, shape.r, shape.s, shape.c), dev, sctx).permute(0, 3, 1, 2)
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
                sctx.keep_alive.append((x, dy, dw))   # read / written on the side stream: their memory must not be handed out again before the join
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                st = _stream()
                if gathered == 2:            # x is the layer input; dw_col[K][256] in the pixel-pair K order, then back to [K][R][S][C]
                    dw_col = _acc_empty((k, 256), dev, sctx)
                    check(lib.dcv_conv2d_wgrad_pairs(ctypes.byref(shape), _ptr(x), _ptr(dy), _ptr(dw_col), pz, st), 'conv2d_wgrad_pairs')
                    check(lib.dcv_pairs_unpack_wgrad(_ptr(dw_col), _ptr(dw), ctypes.byref(shape), st), 'pairs_unpack_wgrad')
                elif gathered:               # x is the layer input; dw_col[K][kpad_g] in the gather kernels' K order, then back to [K][R][S][C]
                    sc = shape.s * shape.c
                    kpad_g = (shape.r * ((sc + 7) // 8 * 8) + 63) // 64 * 64
                    dw_col = _acc_empty((k, kpad_g), dev, sctx)
                    check(lib.dcv_conv2d_wgrad_gather(ctypes.byref(shape), _ptr(x), _ptr(dy), _ptr(dw_col), kpad_g, pz, st), 'conv2d_wgrad_gather')
                    check(lib.dcv_gather_unpack_wgrad(_ptr(dw_col), _ptr(dw), k, shape.r, sc, kpad_g, st), 'gather_unpack_wgrad')
                elif gemm_shape is not None:   # x is the saved im2col matrix: dw_col[K][kpad] = dy^T @ col, then drop the zero padding
                    dw_col = _acc_empty((k, gemm_shape.c), dev, sctx)
                    check(lib.dcv_conv2d_wgrad(ctypes.byref(gemm_shape), _ptr(x), _ptr(dy), _ptr(dw_col), None, dt, algo, pz, st), 'conv2d_wgrad(im2col)')
                    check(lib.dcv_copy_channels_out(_ptr(dw_col), _ptr(dw), k, gemm_shape.c, 0, shape.r * shape.s * shape.c, DCV_F32, st), 'copy_channels_out(dw)')
                else:
                    ws_bytes = int(lib.dcv_conv2d_wgrad_workspace(ctypes.byref(shape), dt, algo))
                    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
                    check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), _ptr(x), _ptr(dy), _ptr(dw), _ptr(ws), dt, algo, pz, st), 'conv2d_wgrad')
        st = _stream()
        if _DEBUG_CAPTURE is not None:
            _DEBUG_CAPTURE.append(dict(wshape=wshape, dz=dz.clone(), y=y.clone(), dy=dy.clone(), pqr=None if pqr is None else pqr.clone(), saved=None if saved is None else saved.clone(), dz_ptr=dz.data_ptr(), y_ptr=y.data_ptr(), dy_ptr=dy.data_ptr(), x=x.clone(), dw=None if dw is None else dw.clone()))
        def ret(name, g):  # gradients written straight into a caller-provided bucket slice are not handed to autograd again
            return None if (g is None or name in targets) else g
        if ctx.notify:
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):   # "enqueued on the current stream": the reducer orders its all-reduce behind it
                _backward_done(grad_out, targets)
        return (dx, ret('weight', dw), ret('bias', dbias) if has_bias else None, ret('bn_w', d_bn_w), ret('bn_b', d_bn_b), ret('gn_w', d_gn_w), ret('gn_b', d_gn_b),
                None, None, None, None, None, None, None, None, None, None, None, None, None, None, None, None)


def _backward_done(grad_out: Optional[dict], targets: dict) -> None:
    """ End of a fused layer's backward: every kernel that writes this layer's parameter gradients has been enqueued on the current stream. Marks the
    bucket slices as written and tells the data-parallel reducer (`flat_params.GradientBucketReducer`), which may now all-reduce a bucket whose last
    producer this was. (Module full-backward hooks fire BEFORE the weight-gradient kernels of a layer whose input needs no gradient are enqueued.) """
    if not grad_out:
        return
    if targets is grad_out:
        grad_out['_written'] = True
    done = grad_out.get('_done')
    if done is not None:
        done()


def conv_block(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: Sequence[int], padding: Sequence[int], dilation: Sequence[int],
               act: int = ACT_NONE, slope: float = 0., norm: Optional[NormConfig] = None, training: bool = True,
               bn_weight=None, bn_bias=None, running_mean=None, running_var=None, num_batches_tracked=None, gn_weight=None, gn_bias=None,
               algo: int = ALGO_AUTO, grad_out: Optional[dict] = None, step_ctx: Optional[StepContext] = None, notify: bool = True, defer_apply: bool = False):
    """ act(conv2d(x, weight) + bias) followed by the configured BatchNorm / GroupNorm, as one autograd node.
    `grad_out` optionally maps 'weight' / 'bias' / 'bn_w' / 'bn_b' / 'gn_w' / 'gn_b' to preallocated fp32 tensors (slices of a
    flat gradient bucket) that backward fills in place instead of returning new tensors. `notify=False`: another node of the same layer (the
    normalisation of a pre-activation block, whose backward runs later) closes the layer's backward (`_backward_done`).
    `defer_apply=True` (the caller promises ONE consumer that takes a `PendingNorm`: a 2x2 average pooling, a residual sum, or `materialize`): a block
    with a normalisation returns its raw output and coefficients instead of running the apply pass. """
    x = as_nhwc(x)
    norm = norm if norm is not None else NormConfig()
    defer_apply = bool(defer_apply and norm.any and _LAZY_APPLY)
    link = {} if defer_apply else None
    out = _ConvBlock.apply(x, weight, bias, bn_weight, bn_bias, gn_weight, gn_bias, running_mean, running_var, num_batches_tracked,
                           tuple(stride), tuple(padding), tuple(dilation), int(act), float(slope), norm, bool(training), int(algo), grad_out, step_ctx, bool(notify), defer_apply, link)
    return PendingNorm(out[0], out[1], link) if defer_apply else out


class PendingNorm:
    """ A tensor-core block's RAW output `y` (conv + bias + activation) and its normalisation coefficients `ab[n][c] = (A, B)`: the one consumer computes
    z = A*y + B inside its own pass — `materialize` (the plain apply pass), a 2x2 / stride-2 average pooling (`avg_pool2d`), a residual sum
    (`link_reduce`). The gradient a consumer returns for `y` is dz, the gradient w.r.t. z: the block's backward turns it into the gradient of y.
    Only the modules of this package ever see one (`DeepcvModule.forward` decides who may). """
    __slots__ = ('y', 'ab', 'link', 'consumed')

    def __init__(self, y: torch.Tensor, ab: torch.Tensor, link: Optional[dict] = None):
        self.y, self.ab, self.link, self.consumed = y, ab, link, False

    shape = property(lambda self: self.y.shape)
    dtype = property(lambda self: self.y.dtype)
    device = property(lambda self: self.y.device)

    def dim(self) -> int:
        return self.y.dim()

    def take(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.consumed:
            raise RuntimeError('deepcv_b200: a pending normalisation was consumed twice (materialize() it before using the tensor in two places)')
        self.consumed = True
        return self.y, self.ab


_PLACEHOLDERS = {}


def _placeholder_grad(shape, dtype, device) -> torch.Tensor:
    """ A tensor of the given shape that occupies one element (stride 0): what `_ApplyNorm.backward` hands to autograd in the slot of y when the real
    gradient travels through the side channel. Never read. """
    key = (dtype, device.type, device.index)
    t = _PLACEHOLDERS.get(key)
    if t is None:
        t = _PLACEHOLDERS[key] = torch.zeros((), dtype=dtype, device=device)
    return t.expand(tuple(shape))


class _ApplyNorm(torch.autograd.Function):
    """ z = A*y + B [+ other] [then 2x2 / stride-2 average pooling] of a `PendingNorm`. Backward hands dz to the producing block (see `PendingNorm`). """

    @staticmethod
    def forward(ctx, y, ab, other, pool: bool, link=None):
        n, c, h, w = y.shape
        ctx.geom, ctx.pool, ctx.has_other, ctx.link = (n, c, h, w), pool, other is not None, link
        st, dt = _stream(), _dt(y)
        fold = link.get('fold') if link is not None else None
        if fold is not None:   # BatchNorm-only block, batch as one image: the coefficients are computed by this very kernel (no finalize launch)
            out = empty_nhwc(n, c, h // 2, w // 2, y.dtype, y.device) if pool else empty_nhwc(n, c, h, w, y.dtype, y.device)
            _bn_apply_fold(fold, y, other, pool, out, n * h, w, c, dt, st)
            return out
        one = ab.dim() == 3 and ab.shape[0] == 1 and n > 1   # BatchNorm-only block: one row of coefficients, the batch is one image of n*h rows (h even when
        ne, he = (1, n * h) if one else (n, h)                # pooled: image boundaries fall on even rows)
        if pool:
            out = empty_nhwc(n, c, h // 2, w // 2, y.dtype, y.device)
            check(lib.dcv_norm_apply_pool_fwd(_ptr(y), _ptr(ab), _ptr(out), ne, he, w, c, dt, st), 'norm_apply_pool_fwd')
        elif other is not None:
            out = empty_nhwc(n, c, h, w, y.dtype, y.device)
            check(lib.dcv_norm_apply_add_fwd(_ptr(y), _ptr(ab), _ptr(other), _ptr(out), ne, he * w, c, dt, st), 'norm_apply_add_fwd')
        else:
            out = empty_nhwc(n, c, h, w, y.dtype, y.device)
            check(lib.dcv_norm_apply_fwd(_ptr(y), _ptr(ab), _ptr(out), ne, he * w, c, dt, st), 'norm_apply_fwd')
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = ctx.geom
        if ctx.pool:
            g = as_nhwc(g.detach())
            if ctx.link is not None and _POOLED_BWD and lib.dcv_norm_bwd_pooled_supported(n, h, w, c, _dt(g)) and g.data_ptr() % 16 == 0:
                ctx.link['pooled_dz'] = g   # the block's normalisation backward reads the pooled gradient in place (no full-resolution dz)
                return _placeholder_grad((n, c, h, w), g.dtype, g.device), None, None, None, None
            dz = empty_nhwc(n, c, h, w, g.dtype, g.device)
            check(lib.dcv_avgpool2d_bwd(_ptr(g), _ptr(dz), n, h, w, c, 2, 2, 2, 2, _dt(g), _stream()), 'avgpool2d_bwd')
            return dz, None, None, None, None
        return g, None, (g if ctx.has_other else None), None, None


def apply_pending(pn: PendingNorm, other: Optional[torch.Tensor] = None, pool: bool = False) -> torch.Tensor:
    y, ab = pn.take()
    if other is not None:
        other = as_nhwc(other, y.dtype)
    return _ApplyNorm.apply(y, ab, other, bool(pool), pn.link)


# ------------------------------------------------------------------------------------------------------------------------------
# Dropout, stand-alone activation and the pre-activation block order `[Dropout] -> norms -> act -> op` (reference: meta/nn.py:535-541, 553)

class DropoutState:
    """ What one `torch.nn.Dropout` of a fused layer needs on the device: a 64-bit seed (drawn from torch's default generator at first use, so
    `torch.manual_seed` controls it) and the call counter (device int32, advanced by a kernel: graph-replay safe). `record`: tests set it to a list
    to receive every mask (uint8, same strides as the input) — the masks the CPU oracle is then given. """

    def __init__(self, device, seed: Optional[int] = None):
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        self.counter = torch.zeros((), dtype=torch.int32, device=device)
        self.record: Optional[list] = None


def _aligned_dense(t: torch.Tensor) -> torch.Tensor:
    t = _dense_like(t)
    return t if t.data_ptr() % 16 == 0 else t.clone()


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p: float, state: DropoutState):
        _require_cuda(x)
        x = _aligned_dense(x)
        y = torch.empty_like(x)
        saved_call = torch.empty((), dtype=torch.int32, device=x.device)
        mask = torch.empty_like(x, dtype=torch.uint8) if state.record is not None else None
        st = _stream()
        check(lib.dcv_dropout(_ptr(x), _ptr(y), _ptr(mask), x.numel(), p, state.seed, _ptr(state.counter), _ptr(saved_call), _dt(x), st), 'dropout')
        check(lib.dcv_counter_add(_ptr(state.counter), 1, st), 'counter_add(dropout call)')
        if mask is not None:
            state.record.append(mask)
        ctx.save_for_backward(saved_call)
        ctx.p, ctx.seed = p, state.seed
        return y

    @staticmethod
    def backward(ctx, dy):
        saved_call, = ctx.saved_tensors
        dy = _aligned_dense(dy.detach())
        dx = torch.empty_like(dy)
        check(lib.dcv_dropout(_ptr(dy), _ptr(dx), None, dy.numel(), ctx.p, ctx.seed, _ptr(saved_call), None, _dt(dy), _stream()), 'dropout(backward)')
        return dx, None, None


def dropout(x: torch.Tensor, p: float, state: DropoutState) -> torch.Tensor:
    """ `torch.nn.Dropout(p)` in training mode: x * keep / (1 - p), keep ~ Bernoulli(1 - p) per element (Philox mask regenerated in backward). """
    if not 0. <= p < 1.:
        raise ValueError(f'dropout probability has to be in [0, 1), but got {p}')
    return x if p == 0. else _Dropout.apply(x, float(p), state)


class _Activation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act: int, slope: float):
        _require_cuda(x)
        x = _aligned_dense(x)
        y = torch.empty_like(x)
        check(lib.dcv_activation_fwd(_ptr(x), _ptr(y), x.numel(), act, slope, _dt(x), _stream()), 'activation_fwd')
        ctx.save_for_backward(y)
        ctx.cfg = (act, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, = ctx.saved_tensors
        dy = _aligned_dense(dy.detach())
        if dy.stride() != y.stride():
            dy = as_nhwc(dy) if (y.dim() == 4 and is_nhwc(y)) else dy.contiguous()
        dx = torch.empty_like(y)
        check(lib.dcv_activation_bwd(_ptr(dy), _ptr(y), _ptr(dx), y.numel(), ctx.cfg[0], ctx.cfg[1], _dt(y), _stream()), 'activation_bwd')
        return dx, None, None


def activation(x: torch.Tensor, act: int, slope: float = 0.) -> torch.Tensor:
    """ Stand-alone ReLU / LeakyReLU / Sigmoid (the activation of a pre-activation block, applied to the block's input). """
    return x if act == ACT_NONE else _Activation.apply(x, int(act), float(slope))


class _PreNormAct(torch.autograd.Function):
    """ act(GroupNorm(BatchNorm(x))) on the block INPUT (pre-activation order): statistics pass, O(N*C) finalize, one fused affine pass, activation in
    place. Backward: du = da * act'(a), then the normalisation adjoint dx = P*du + Q*x + R (reduce -> finalize -> apply). """

    @staticmethod
    def forward(ctx, x, bn_w, bn_b, gn_w, gn_b, rm, rv, nbt, act, slope, cfg: NormConfig, training: bool, grad_out, sctx):
        _require_cuda(x)
        n, c, h, w = x.shape
        dev, st, dt, pz = x.device, _stream(), _dt(x), _pz(sctx)
        stats = _acc_empty((n, c, 2), dev, sctx)
        check(lib.dcv_norm_stats(_ptr(x), _ptr(stats), n, h * w, c, dt, pz, st), 'norm_stats')
        groups = cfg.gn_groups if cfg.use_gn else 1
        saved = torch.empty((int(lib.dcv_norm_saved_floats(n, c, groups)),), dtype=torch.float32, device=dev)
        ab = torch.empty((n, c, 2), dtype=torch.float32, device=dev)
        prm = _norm_params(cfg, n, c, h * w, training, bn_w, bn_b, rm, rv, nbt, gn_w, gn_b)
        check(lib.dcv_norm_fwd_finalize(ctypes.byref(prm), _ptr(stats), _ptr(ab), _ptr(saved), st), 'norm_fwd_finalize')
        a = empty_nhwc(n, c, h, w, x.dtype, dev)
        check(lib.dcv_norm_apply_fwd(_ptr(x), _ptr(ab), _ptr(a), n, h * w, c, dt, st), 'norm_apply_fwd')
        if act != ACT_NONE:
            check(lib.dcv_activation_fwd(_ptr(a), _ptr(a), a.numel(), act, slope, dt, st), 'activation_fwd')
        ctx.save_for_backward(x, a, stats, saved, bn_w, bn_b, gn_w, gn_b, rm, rv)
        ctx.cfg = (act, slope, cfg, training, grad_out, sctx)
        return a

    @staticmethod
    def backward(ctx, da):
        x, a, stats, saved, bn_w, bn_b, gn_w, gn_b, rm, rv = ctx.saved_tensors
        act, slope, cfg, training, grad_out, sctx = ctx.cfg
        n, c, h, w = x.shape
        dev, st, dt, pz = x.device, _stream(), _dt(x), _pz(sctx)
        da = as_nhwc(da.detach(), x.dtype)
        targets = grad_out if (grad_out and not grad_out.get('_written', False)) else {}
        du = da
        if act != ACT_NONE:
            du = empty_nhwc(n, c, h, w, x.dtype, dev)
            check(lib.dcv_activation_bwd(_ptr(da), _ptr(a), _ptr(du), a.numel(), act, slope, dt, st), 'activation_bwd')
        s_nc = _acc_empty((n, c, 2), dev, sctx)
        check(lib.dcv_norm_bwd_reduce(_ptr(du), _ptr(x), _ptr(s_nc), n, h * w, c, dt, pz, st), 'norm_bwd_reduce')
        f32 = dict(dtype=torch.float32, device=dev)
        pqr = torch.empty((n, c, 3), **f32)

        def grad_buf(name, param):
            if param is None:
                return None
            t = targets.get(name)
            return torch.empty((c,), **f32) if t is None else t
        d_bn_w, d_bn_b = (grad_buf('bn_w', bn_w), grad_buf('bn_b', bn_b)) if cfg.use_bn else (None, None)
        d_gn_w, d_gn_b = (grad_buf('gn_w', gn_w), grad_buf('gn_b', gn_b)) if cfg.use_gn else (None, None)
        prm = _norm_params(cfg, n, c, h * w, training, bn_w, bn_b, rm, rv, None, gn_w, gn_b)
        check(lib.dcv_norm_bwd_finalize(ctypes.byref(prm), _ptr(stats), _ptr(s_nc), _ptr(saved), _ptr(pqr), _ptr(d_bn_w), _ptr(d_bn_b), _ptr(d_gn_w), _ptr(d_gn_b), st), 'norm_bwd_finalize')
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty_nhwc(n, c, h, w, x.dtype, dev)
            check(lib.dcv_act_norm_bwd_apply(_ptr(du), _ptr(x), _ptr(pqr), _ptr(dx), None, ACT_NONE, 0., n, h * w, c, dt, pz, st), 'act_norm_bwd_apply')

        def ret(name, g):
            return None if (g is None or name in targets) else g
        _backward_done(grad_out, targets)
        return dx, ret('bn_w', d_bn_w), ret('bn_b', d_bn_b), ret('gn_w', d_gn_w), ret('gn_b', d_gn_b), None, None, None, None, None, None, None, None, None


def pre_norm_act(x: torch.Tensor, norm: NormConfig, training: bool = True, act: int = ACT_NONE, slope: float = 0., bn_weight=None, bn_bias=None, running_mean=None,
                 running_var=None, num_batches_tracked=None, gn_weight=None, gn_bias=None, grad_out: Optional[dict] = None, step_ctx: Optional[StepContext] = None) -> torch.Tensor:
    """ The head of a pre-activation block (reference meta/nn.py:553 `(*norm_ops, act_fn(), layer_op)`): act(norms(x)) as one autograd node. It closes the
    layer's backward (its backward runs after the convolution's): pass `notify=False` to the `conv_block` that follows. """
    if x.dim() != 4:
        raise NotImplementedError('deepcv_b200: normalisation in front of a fully connected layer is not built (image tensors only)')
    return _PreNormAct.apply(as_nhwc(x), bn_weight, bn_bias, gn_weight, gn_bias, running_mean, running_var, num_batches_tracked, int(act), float(slope), norm, bool(training),
                             grad_out, step_ctx)


# ------------------------------------------------------------------------------------------------------------------------------
# Few-channel convolution blocks (csrc/conv_small.cu): raw outputs with a PENDING normalisation travel between submodules

_FUSE_SC_BWD = os.environ.get('DCV_NO_SC_FUSED_BWD') is None   # A/B switch: separate weight-gradient and data-gradient launches for the few-channel layers
_USE_SC = os.environ.get('DCV_NO_SC') is None   # tuning aid: DCV_NO_SC=1 keeps the few-channel layers on the direct kernels + separate normalisation passes


class _NormLink:
    """ Device-side state of one block's pending BatchNorm o GroupNorm for ONE step: the `dcv_sc_norm` descriptor and the buffers it points to (kept
    alive here), shared by the producer's and the consumer's autograd nodes. """

    def __init__(self, n, c, hw, cfg: NormConfig, training, bn_w, bn_b, rm, rv, nbt, gn_w, gn_b, device, sctx: Optional[StepContext], need_backward: bool):
        f = lambda which: int(lib.dcv_sc_norm_floats(n, c, which))
        self.stats = torch.empty(f(0), dtype=torch.float32, device=device)
        bn_training = bool(training or rm is None or rv is None)
        self.bn_sums = _zeroed_acc((f(1),), device, sctx) if (cfg.use_bn and bn_training) else None
        self.s_nc = None   # diagnostic output of the kernels (sum dz, sum dz*y per (image, channel)); nothing reads it
        self.u_sums = _zeroed_acc((f(3),), device, sctx) if need_backward else None
        self.coef = torch.empty(f(4), dtype=torch.float32, device=device) if need_backward else None   # forward coefficients cached for the backward kernels
        self.d_nc = torch.empty(f(5), dtype=torch.float32, device=device) if need_backward else None
        self.keep = (bn_w, bn_b, rm, rv, nbt, gn_w, gn_b)
        self.struct = ScNorm(1, n, c, hw, int(cfg.use_bn), int(bn_training), cfg.bn_eps, cfg.bn_momentum, _ptr(bn_w), _ptr(bn_b), _ptr(rm), _ptr(rv), _ptr(nbt),
                             int(cfg.use_gn), cfg.gn_groups, cfg.gn_eps, _ptr(gn_w), _ptr(gn_b), _ptr(self.stats), _ptr(self.bn_sums), _ptr(self.s_nc), _ptr(self.u_sums), _ptr(self.coef), _ptr(self.d_nc))
        self.update_running = bool(cfg.use_bn and training and rm is not None and rv is not None)
        self.consumed = False

    def ref(self):
        return ctypes.byref(self.struct)


def _zeroed_acc(shape, device, sctx: Optional[StepContext]) -> torch.Tensor:
    """ An accumulator that is zero when its first kernel runs: a slice of the step's pre-zeroed arena, or a fresh buffer + one memset. """
    t = _acc_empty(shape, device, sctx)
    if not _pz(sctx):
        check(lib.dcv_fill_zero(_ptr(t), t.numel() * 4, _stream()), 'fill_zero(accumulator)')
    return t


class PendingAffine:
    """ A block's RAW output `y` (conv + bias + activation) whose BatchNorm o GroupNorm has not been applied: the consumer applies z = A[n][c]*y + B[n][c]
    while loading (few-channel convolution, average pooling) or `materialize()`s it. Exactly one consumer: its backward produces the producer's sums. Not
    a tensor on purpose: only the modules of this package (`FusedLayer`, `AvgPool2d`, `DeepcvModule.forward`) ever see one. """
    __slots__ = ('y', 'link')

    def __init__(self, y: torch.Tensor, link: _NormLink):
        self.y, self.link = y, link

    shape = property(lambda self: self.y.shape)
    dtype = property(lambda self: self.y.dtype)
    device = property(lambda self: self.y.device)

    def take(self) -> Tuple[torch.Tensor, _NormLink]:
        if self.link.consumed:
            raise RuntimeError('deepcv_b200: a pending normalisation was consumed twice (materialize() it before using the tensor in two places)')
        self.link.consumed = True
        return self.y, self.link


def sc_conv_supported(x_shape, weight: torch.Tensor, stride, padding, dilation, dtype: torch.dtype) -> bool:
    if not _USE_SC or dtype != torch.bfloat16:
        return False
    n, c, h, w = x_shape
    k, cw, r, s = weight.shape
    if cw != c:
        return False
    p = (h + 2 * padding[0] - dilation[0] * (r - 1) - 1) // stride[0] + 1
    q = (w + 2 * padding[1] - dilation[1] * (s - 1) - 1) // stride[1] + 1
    shape = ConvShape(n, h, w, c, k, r, s, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], p, q)
    return bool(lib.dcv_sc_conv_supported(ctypes.byref(shape), DCV_BF16))


class _ScConv(torch.autograd.Function):
    """ One few-channel block: forward = ONE launch (normalise-on-load of the input, convolution, bias, activation, statistics), backward = weight-gradient
    launch + data-gradient launch. `x` is a plain tensor or a producer's raw output (meta['xlink'] set); the returned y is raw when the block has a
    normalisation (the caller wraps it into a `PendingAffine`). The "gradient" exchanged with a producer for its raw output is dz, the gradient w.r.t. the
    NORMALISED tensor — a private protocol between the nodes of this path, which is why raw outputs never leave it. """

    @staticmethod
    def forward(ctx, x, weight, bias, bn_w, bn_b, gn_w, gn_b, meta):
        shape, xlink, cfg, sctx = meta['shape'], meta['xlink'], meta['cfg'], meta['sctx']
        n, k, p, q = shape.n, shape.k, shape.p, shape.q
        dev, st = x.device, _stream()
        w_op = _weight_operand(weight, torch.bfloat16, sctx)
        y = empty_nhwc(n, k, p, q, torch.bfloat16, dev)
        ylink = None
        if cfg.any:
            ylink = _NormLink(n, k, p * q, cfg, meta['training'], bn_w, bn_b, meta['rm'], meta['rv'], meta['nbt'], gn_w, gn_b, dev, sctx, need_backward=meta['need_backward'])
        check(lib.dcv_sc_conv_fwd(ctypes.byref(shape), _ptr(x), xlink.ref() if xlink else None, int(bool(xlink and xlink.update_running)), _ptr(w_op), _ptr(bias),
                                  meta['act'], meta['slope'], _ptr(y), ylink.ref() if ylink else None, st), 'sc_conv_fwd')
        ctx.save_for_backward(x, w_op, y)
        ctx.meta, ctx.ylink, ctx.has_bias = meta, ylink, bias is not None
        ctx.has_affine = (bn_w is not None, bn_b is not None, gn_w is not None, gn_b is not None)
        meta['ylink'] = ylink
        return y

    @staticmethod
    def backward(ctx, dz):
        x, w_op, y = ctx.saved_tensors
        meta, ylink = ctx.meta, ctx.ylink
        shape, xlink, cfg, sctx, grad_out = meta['shape'], meta['xlink'], meta['cfg'], meta['sctx'], meta['grad_out']
        dev, st = y.device, _stream()
        dz = as_nhwc(dz.detach(), torch.bfloat16)
        targets = grad_out if (grad_out and not grad_out.get('_written', False)) else {}

        def target(name, numel_shape, wanted):   # accumulated into by the kernel: a (zeroed) bucket slice, or a fresh zeroed buffer handed to autograd
            if not wanted:
                return None
            t = targets.get(name)
            return t if t is not None else _zeroed_acc(numel_shape, dev, sctx if name in ('weight', 'bias') else None)
        k, c = shape.k, shape.c
        dw = target('weight', (k, shape.r, shape.s, c), ctx.needs_input_grad[1])
        if dw is not None and 'weight' not in targets:
            dw = dw.permute(0, 3, 1, 2)
        dbias = target('bias', (k,), ctx.has_bias and ctx.needs_input_grad[2])
        # normalisation parameter gradients are OVERWRITTEN by the kernel
        def plain(name, on):
            if not on:
                return None
            t = targets.get(name)
            return t if t is not None else torch.empty((k,), dtype=torch.float32, device=dev)
        d_bn_w, d_bn_b = plain('bn_w', cfg.use_bn and ctx.has_affine[0]), plain('bn_b', cfg.use_bn and ctx.has_affine[1])
        d_gn_w, d_gn_b = plain('gn_w', cfg.use_gn and ctx.has_affine[2]), plain('gn_b', cfg.use_gn and ctx.has_affine[3])
        dx = None
        if dw is not None and ctx.needs_input_grad[0] and _FUSE_SC_BWD and lib.dcv_sc_conv_bwd_supported(ctypes.byref(shape), DCV_BF16):
            # one launch for both gradients: the dy tile is staged once (csrc/conv_small.cu, sc_bwd_kernel)
            dx = empty_nhwc(shape.n, c, shape.h, shape.w, torch.bfloat16, dev)
            check(lib.dcv_sc_conv_bwd(ctypes.byref(shape), _ptr(x), xlink.ref() if xlink else None, _ptr(dz), _ptr(y), ylink.ref() if ylink else None, meta['act'], meta['slope'],
                                      _ptr(w_op), _ptr(dx), _ptr(dw), _ptr(dbias), _ptr(d_bn_w), _ptr(d_bn_b), _ptr(d_gn_w), _ptr(d_gn_b), st), 'sc_conv_bwd')
            _backward_done(grad_out, targets)
            return (dx, None if 'weight' in targets else dw, None if (dbias is None or 'bias' in targets) else dbias, None if (d_bn_w is None or 'bn_w' in targets) else d_bn_w,
                    None if (d_bn_b is None or 'bn_b' in targets) else d_bn_b, None if (d_gn_w is None or 'gn_w' in targets) else d_gn_w, None if (d_gn_b is None or 'gn_b' in targets) else d_gn_b, None)
        if dw is not None:
            check(lib.dcv_sc_conv_wgrad(ctypes.byref(shape), _ptr(x), xlink.ref() if xlink else None, _ptr(dz), _ptr(y), ylink.ref() if ylink else None, meta['act'], meta['slope'],
                                        _ptr(dw), _ptr(dbias), _ptr(d_bn_w), _ptr(d_bn_b), _ptr(d_gn_w), _ptr(d_gn_b), st), 'sc_conv_wgrad')
        if ctx.needs_input_grad[0]:
            dx = empty_nhwc(shape.n, c, shape.h, shape.w, torch.bfloat16, dev)
            check(lib.dcv_sc_conv_dgrad(ctypes.byref(shape), _ptr(dz), _ptr(y), ylink.ref() if ylink else None, meta['act'], meta['slope'], _ptr(w_op), _ptr(dx),
                                        _ptr(x) if xlink else None, xlink.ref() if xlink else None, st), 'sc_conv_dgrad')

        def ret(name, g):
            return None if (g is None or name in targets) else g
        _backward_done(grad_out, targets)
        return dx, ret('weight', dw), ret('bias', dbias), ret('bn_w', d_bn_w), ret('bn_b', d_bn_b), ret('gn_w', d_gn_w), ret('gn_b', d_gn_b), None


def sc_conv_block(x, weight: torch.Tensor, bias: Optional[torch.Tensor], padding: Sequence[int], act: int = ACT_NONE, slope: float = 0., norm: Optional[NormConfig] = None,
                  training: bool = True, bn_weight=None, bn_bias=None, running_mean=None, running_var=None, num_batches_tracked=None, gn_weight=None, gn_bias=None,
                  grad_out: Optional[dict] = None, step_ctx: Optional[StepContext] = None):
    """ A few-channel block (`sc_conv_supported`) on the fused kernels. `x`: bf16 NHWC tensor or `PendingAffine`. Returns a `PendingAffine` when the block has
    a normalisation, else the activation tensor. """
    xlink = None
    if isinstance(x, PendingAffine):
        x, xlink = x.take()
    else:
        x = as_nhwc(x)
    norm = norm if norm is not None else NormConfig()
    n, c, h, w = x.shape
    k, _, r, s = weight.shape
    shape = ConvShape(n, h, w, c, k, r, s, 1, 1, padding[0], padding[1], 1, 1, h, w)
    meta = dict(shape=shape, xlink=xlink, cfg=norm, sctx=step_ctx, act=int(act), slope=float(slope), training=bool(training), rm=running_mean, rv=running_var,
                nbt=num_batches_tracked, grad_out=grad_out, ylink=None, need_backward=torch.is_grad_enabled())
    y = _ScConv.apply(x, weight, bias, bn_weight, bn_bias, gn_weight, gn_bias, meta)
    return PendingAffine(y, meta['ylink']) if meta['ylink'] is not None else y


class _ScAffinePool(torch.autograd.Function):
    """ z = A*avgpool(y) + B (pool = 1: plain materialisation of a pending normalisation). Backward: dz at full resolution + the producer's sums. """

    @staticmethod
    def forward(ctx, y, link: _NormLink, pool: int):
        n, c, h, w = y.shape
        z = empty_nhwc(n, c, h // pool, w // pool, y.dtype, y.device)
        check(lib.dcv_sc_affine_pool_fwd(_ptr(y), link.ref(), int(link.update_running), _ptr(z), n, h, w, c, pool, _stream()), 'sc_affine_pool_fwd')
        ctx.save_for_backward(y)
        ctx.link, ctx.pool = link, pool
        return z

    @staticmethod
    def backward(ctx, dzp):
        y, = ctx.saved_tensors
        n, c, h, w = y.shape
        dzp = as_nhwc(dzp.detach(), y.dtype)
        dz = empty_nhwc(n, c, h, w, y.dtype, y.device)
        check(lib.dcv_sc_affine_pool_bwd(_ptr(dzp), _ptr(y), ctx.link.ref(), _ptr(dz), n, h, w, c, ctx.pool, _stream()), 'sc_affine_pool_bwd')
        return dz, None, None


def sc_affine_pool(pa: PendingAffine, pool: int) -> torch.Tensor:
    y, link = pa.take()
    return _ScAffinePool.apply(y, link, int(pool))


class PendingFlatten:
    """ An image tensor (NHWC memory) whose `torch.nn.Flatten` has not been materialised: the fully connected layer that follows reads the NHWC tensor
    through the (C, H, W)-order index map inside its kernels (`dcv_linear_fwd(..., x_nhwc_channels)`), so the transposed copy — and its mirror in backward
    — is never made. Only handed to a consumer that declared `accepts_pending_flatten` (`DeepcvModule.forward` decides); `materialize()` otherwise. """
    __slots__ = ('x',)

    def __init__(self, x: torch.Tensor):
        self.x = x

    shape = property(lambda self: torch.Size((self.x.shape[0], self.x.shape[1] * self.x.shape[2] * self.x.shape[3])))
    dtype = property(lambda self: self.x.dtype)
    device = property(lambda self: self.x.device)


def materialize(x):
    """ `PendingAffine` / `PendingNorm` -> the normalised tensor; `PendingFlatten` -> the flattened (N, C*H*W) tensor; tensors pass through. """
    if isinstance(x, PendingAffine):
        return sc_affine_pool(x, 1)
    if isinstance(x, PendingNorm):
        return apply_pending(x)
    if isinstance(x, PendingFlatten):
        return flatten_nchw(x.x)
    return x


# ------------------------------------------------------------------------------------------------------------------------------
# Average pooling (reference: meta/submodule_creators.py:163-176 -> torch.nn.AvgPool2d)

class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, stride):
        _require_cuda(x)
        n, c, h, w = x.shape
        if h < kernel[0] or w < kernel[1]:
            raise RuntimeError(f'deepcv_b200: average pooling kernel {kernel} larger than input {h}x{w}')
        p, q = (h - kernel[0]) // stride[0] + 1, (w - kernel[1]) // stride[1] + 1
        y = empty_nhwc(n, c, p, q, x.dtype, x.device)
        check(lib.dcv_avgpool2d_fwd(_ptr(x), _ptr(y), n, h, w, c, kernel[0], kernel[1], stride[0], stride[1], _dt(x), _stream()), 'avgpool2d_fwd')
        ctx.geom = (n, c, h, w, kernel, stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, kernel, stride = ctx.geom
        dy = as_nhwc(dy.detach())
        dx = empty_nhwc(n, c, h, w, dy.dtype, dy.device)
        check(lib.dcv_avgpool2d_bwd(_ptr(dy), _ptr(dx), n, h, w, c, kernel[0], kernel[1], stride[0], stride[1], _dt(dy), _stream()), 'avgpool2d_bwd')
        return dx, None, None


def avg_pool2d(x: torch.Tensor, kernel_size, stride=None) -> torch.Tensor:
    kernel = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    stride = kernel if stride is None else ((stride, stride) if isinstance(stride, int) else tuple(stride))
    if isinstance(x, PendingNorm):
        if kernel == (2, 2) and stride == (2, 2) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            return apply_pending(x, pool=True)   # the normalised full-resolution tensor is never written
        x = materialize(x)
    return _AvgPool.apply(as_nhwc(x), kernel, stride)


# ------------------------------------------------------------------------------------------------------------------------------
# Links (reference: meta/submodule_creators.py:43-65, 272-332; meta/nn.py:665-676)

class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow, align_corners):
        _require_cuda(x)
        n, c, h, w = x.shape
        y = empty_nhwc(n, c, oh, ow, x.dtype, x.device)
        check(lib.dcv_bilinear_fwd(_ptr(x), _ptr(y), n, h, w, c, oh, ow, int(align_corners), _dt(x), _stream()), 'bilinear_fwd')
        ctx.geom = (n, c, h, w, oh, ow, int(align_corners))
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, oh, ow, align = ctx.geom
        dy = as_nhwc(dy.detach())
        dx32 = empty_nhwc(n, c, h, w, torch.float32, dy.device)
        check(lib.dcv_bilinear_bwd(_ptr(dy), _ptr(dx32), n, h, w, c, oh, ow, align, _dt(dy), _stream()), 'bilinear_bwd')
        return _cast_raw(dx32, dy.dtype), None, None, None


def bilinear_resize(x: torch.Tensor, size: Sequence[int], align_corners: bool = False) -> torch.Tensor:
    """ `F.interpolate(x, size, mode='bilinear', align_corners)`. An exact 2x reduction with align_corners=False samples at
    2d+0.5, i.e. it *is* the 2x2 average: that case runs the vectorised pooling kernels. """
    x = as_nhwc(x)
    oh, ow = int(size[0]), int(size[1])
    if not align_corners and x.shape[2] == 2 * oh and x.shape[3] == 2 * ow:
        return _AvgPool.apply(x, (2, 2), (2, 2))
    return _Bilinear.apply(x, oh, ow, bool(align_corners))


class _Axpby(torch.autograd.Function):
    """ out = alpha * (a + b) : residual sum (alpha=1) and the two-operand mean (alpha=0.5). """

    @staticmethod
    def forward(ctx, a, b, alpha):
        _require_cuda(a, b)
        out = torch.empty_like(a)
        check(lib.dcv_axpby(_ptr(a), _ptr(b), _ptr(out), alpha, alpha, a.numel(), _dt(a), _stream()), 'axpby')
        ctx.alpha = alpha
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.alpha == 1.:
            return g, g, None
        g = _dense_like(g.detach())
        ga = torch.empty_like(g)
        check(lib.dcv_axpby(_ptr(g), None, _ptr(ga), ctx.alpha, 0., g.numel(), _dt(g), _stream()), 'axpby')
        return ga, ga, None


class _Concat(torch.autograd.Function):
    """ Channel concatenation of NHWC tensors in one launch each way (`dcv_link_concat_*`); `pools[i] == 2` marks a tensor twice the output's spatial size
    that enters as its 2x2 average (a `dense_link` reference rescaled by exactly 1/2). More than `LINK_MAX_SOURCES` same-size tensors: one slice copy each. """

    @staticmethod
    def forward(ctx, pools, *tensors):
        _require_cuda(*tensors)
        n, _, h, w = tensors[0].shape
        h, w = h // pools[0], w // pools[0]
        channels = [t.shape[1] for t in tensors]
        out = empty_nhwc(n, sum(channels), h, w, tensors[0].dtype, tensors[0].device)
        ctx.channels, ctx.pools, ctx.fused = channels, pools, len(tensors) <= LINK_MAX_SOURCES
        if ctx.fused:
            srcs = (LinkSource * len(tensors))(*[LinkSource(_ptr(t), c, p) for t, c, p in zip(tensors, channels, pools)])
            check(lib.dcv_link_concat_fwd(srcs, len(tensors), _ptr(out), n, h, w, _dt(out), _stream()), 'link_concat_fwd')
            return out
        off = 0
        for t, c in zip(tensors, channels):
            check(lib.dcv_copy_channels_in(_ptr(t), _ptr(out), n * h * w, c, sum(channels), off, _dt(t), _stream()), 'copy_channels_in')
            off += c
        return out

    @staticmethod
    def backward(ctx, g):
        g = as_nhwc(g.detach())
        n, ctot, h, w = g.shape
        grads = [empty_nhwc(n, c, h * p, w * p, g.dtype, g.device) if ctx.needs_input_grad[i + 1] else None for i, (c, p) in enumerate(zip(ctx.channels, ctx.pools))]
        if ctx.fused:
            dsts = (LinkSource * len(grads))(*[LinkSource(_ptr(gi) if gi is not None else None, c, p) for gi, c, p in zip(grads, ctx.channels, ctx.pools)])
            check(lib.dcv_link_concat_bwd(_ptr(g), dsts, len(grads), n, h, w, _dt(g), _stream()), 'link_concat_bwd')
            return (None, *grads)
        off = 0
        for gi, c in zip(grads, ctx.channels):
            if gi is not None:
                check(lib.dcv_copy_channels_out(_ptr(g), _ptr(gi), n * h * w, ctot, off, c, _dt(g), _stream()), 'copy_channels_out')
            off += c
        return (None, *grads)


def link_concat_rescaled(tensors: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """ `dense_link` body when every `_from` tensor either has the first tensor's spatial shape or exactly twice it (bilinear, align_corners=False: the
    2x2 average): rescale + concat in one launch. None when the list is not of that form (the caller rescales, then `link_reduce`s). """
    if len(tensors) < 2 or len(tensors) > LINK_MAX_SOURCES or any(not (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 4) for t in tensors):
        return None
    n, _, h, w = tensors[0].shape
    pools = []
    for t in tensors:
        if t.shape[0] != n or tuple(t.shape[2:]) not in ((h, w), (2 * h, 2 * w)):
            return None
        pools.append(1 if tuple(t.shape[2:]) == (h, w) else 2)
    dtype = tensors[0].dtype
    return _Concat.apply(tuple(pools), *[as_nhwc(t, dtype) for t in tensors])


def link_reduce(tensors: List[torch.Tensor], reduction: str) -> torch.Tensor:
    """ 'sum' / 'mean' (elementwise over the list) or 'concat' (channel dim, first tensor first). """
    if len(tensors) == 1 and reduction in ('sum', 'mean', 'concat'):
        return materialize(tensors[0])
    if isinstance(tensors[0], PendingNorm):   # a residual link right behind a block: the sum rides in the block's apply pass
        if reduction == 'sum' and len(tensors) == 2 and tuple(tensors[1].shape) == tuple(tensors[0].shape) and tensors[1].dtype == tensors[0].dtype:
            return apply_pending(tensors[0], other=tensors[1])
        tensors = [materialize(tensors[0])] + list(tensors[1:])
    dtype = tensors[0].dtype
    tensors = [as_nhwc(t, dtype) for t in tensors]
    if reduction == 'concat':
        if any(t.shape[0] != tensors[0].shape[0] or t.shape[2:] != tensors[0].shape[2:] for t in tensors):
            raise RuntimeError(f'deepcv_b200: cannot concatenate tensors of shapes {[tuple(t.shape) for t in tensors]} on the channel dim')
        return _Concat.apply((1,) * len(tensors), *tensors)
    if reduction in ('sum', 'mean'):
        if any(t.shape != tensors[0].shape for t in tensors):
            raise RuntimeError(f'deepcv_b200: cannot {reduction} tensors of different shapes {[tuple(t.shape) for t in tensors]}')
        acc = tensors[0]
        for t in tensors[1:]:
            acc = _Axpby.apply(acc, t, 1.)
        if reduction == 'mean':
            acc = _Scale.apply(acc, 1. / len(tensors))
        return acc
    raise ValueError(f'Error: Invalid "{reduction}" reduction function name. Valid reduction functions are "mean", "sum", "concat" and "none".')


class _Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, alpha):
        out = torch.empty_like(a)
        check(lib.dcv_axpby(_ptr(a), None, _ptr(out), alpha, 0., a.numel(), _dt(a), _stream()), 'axpby')
        ctx.alpha = alpha
        return out

    @staticmethod
    def backward(ctx, g):
        g = _dense_like(g.detach())
        out = torch.empty_like(g)
        check(lib.dcv_axpby(_ptr(g), None, _ptr(out), ctx.alpha, 0., g.numel(), _dt(g), _stream()), 'axpby')
        return out, None


class _Fork(torch.autograd.Function):
    """ Two aliases of one tensor whose gradients are summed by `dcv_axpby` (a tensor kept for a later link has two consumers;
    without this node autograd would add the two gradients with an ATen kernel). """

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None or g2 is None:
            return g1 if g2 is None else g2
        g1, g2 = _dense_like(g1.detach()), _dense_like(g2.detach())
        if g1.stride() != g2.stride() or g1.dtype != g2.dtype:
            g2 = as_nhwc(g2, g1.dtype) if g1.dim() == 4 else _cast_raw(g2.contiguous(), g1.dtype)
            g1 = as_nhwc(g1) if g1.dim() == 4 else g1.contiguous()
        out = torch.empty_like(g1)
        check(lib.dcv_axpby(_ptr(g1), _ptr(g2), _ptr(out), 1., 1., g1.numel(), _dt(g1), _stream()), 'axpby')
        return out


def fork(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if not (x.is_cuda and x.requires_grad and torch.is_grad_enabled()):
        return x, x
    return _Fork.apply(x)


# ------------------------------------------------------------------------------------------------------------------------------
# Flatten + fully connected head (reference: conf/base/parameters.yml:87-88; meta/submodule_creators.py:268-269)

class _FlattenNCHW(torch.autograd.Function):
    """ torch.nn.Flatten on the logical N x C x H x W tensor: (C, H, W)-major features from NHWC memory. """

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        n, c, h, w = x.shape
        out = torch.empty((n, c * h * w), dtype=x.dtype, device=x.device)
        check(lib.dcv_nhwc_to_nchw(_ptr(x), _dt(x), _ptr(out), _dt(out), n, c, h, w, _stream()), 'nhwc_to_nchw')
        ctx.shape = (n, c, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = ctx.shape
        g = g.detach()
        g = g if g.is_contiguous() else g.contiguous()
        dx = empty_nhwc(n, c, h, w, g.dtype, g.device)
        check(lib.dcv_nchw_to_nhwc(_ptr(g), _dt(g), _ptr(dx), _dt(dx), n, c, h, w, _stream()), 'nchw_to_nhwc')
        return dx


def flatten_nchw(x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 4:
        return x.flatten(1)  # a view: nothing to compute
    if x.shape[2] == 1 and x.shape[3] == 1 and is_nhwc(x):
        return x.reshape(x.shape[0], x.shape[1])  # NHWC with H=W=1 is already (N, C): a view
    return _FlattenNCHW.apply(as_nhwc(x))


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, slope, grad_out, sctx, nhwc_c=0):
        _require_cuda(x, weight)
        ctx.x_shape = tuple(x.shape)
        m, k = (x.shape[0], x.shape[1] * x.shape[2] * x.shape[3]) if nhwc_c else x.shape   # nhwc_c > 0: x is the NHWC image tensor, Flatten fused into the kernels
        n = weight.shape[0]
        if weight.shape[1] != k:
            raise RuntimeError(f'deepcv_b200: linear layer expects {weight.shape[1]} input features, got {k}')
        w = weight if weight.is_contiguous() else weight.contiguous()
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        check(lib.dcv_linear_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), m, n, k, act, slope, _dt(x), DCV_F32, nhwc_c, _stream()), 'linear_fwd')
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (act, slope, bias is not None, grad_out, sctx, nhwc_c)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        act, slope, has_bias, grad_out, sctx, nhwc_c = ctx.cfg
        targets = grad_out if (grad_out and not grad_out.get('_written', False)) else {}
        m, k = (x.shape[0], x.shape[1] * x.shape[2] * x.shape[3]) if nhwc_c else x.shape
        n = w.shape[0]
        dy = _cast_raw(dy.detach().contiguous(), torch.float32)
        f32 = dict(dtype=torch.float32, device=x.device)
        dx = None
        if ctx.needs_input_grad[0]:   # fused Flatten: the gradient comes back in the NHWC layout of the image tensor
            dx = empty_nhwc(*x.shape, x.dtype, x.device) if nhwc_c else torch.empty((m, k), dtype=x.dtype, device=x.device)
        dw = targets.get('weight', None) if ctx.needs_input_grad[1] else None
        if dw is None and ctx.needs_input_grad[1]:
            dw = _acc_empty((n, k), x.device, sctx)
        db = None
        if has_bias:
            db = targets.get('bias', None)
            db = _acc_empty((n,), x.device, sctx) if db is None else db
        dpre = torch.empty((m, n), **f32)
        check(lib.dcv_linear_bwd(_ptr(x), _ptr(w), _ptr(y), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), _ptr(dpre), m, n, k, act, slope, _dt(x), DCV_F32, _pz(sctx), nhwc_c, _stream()), 'linear_bwd')
        _backward_done(grad_out, targets)
        return dx, (None if 'weight' in targets else dw), (None if ('bias' in targets or not has_bias) else db), None, None, None, None, None


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE, slope: float = 0., grad_out: Optional[dict] = None,
               step_ctx: Optional[StepContext] = None) -> torch.Tensor:
    """ act(x @ weight.T + bias) with fp32 output (logits / losses stay fp32 whatever the activation dtype). `x`: (N, K) tensor, image tensor (flattened in
    (C, H, W) order first) or `PendingFlatten` (the same, fused into the layer's kernels when the head is small enough). """
    if isinstance(x, PendingFlatten):
        t = x.x
        if t.dim() == 4 and is_nhwc(t) and t.shape[2] * t.shape[3] > 1 and lib.dcv_linear_flatten_fused(int(weight.shape[0]), int(t.shape[1] * t.shape[2] * t.shape[3])):
            if weight.shape[1] != t.shape[1] * t.shape[2] * t.shape[3]:
                raise RuntimeError(f'deepcv_b200: linear layer expects {weight.shape[1]} input features, got {t.shape[1] * t.shape[2] * t.shape[3]}')
            return _Linear.apply(t, weight, bias, int(act), float(slope), grad_out, step_ctx, int(t.shape[1]))
        x = flatten_nchw(t)
    if x.dim() != 2:
        x = flatten_nchw(x) if x.dim() == 4 else x.reshape(x.shape[0], -1)
    if not x.is_contiguous():
        x = x.contiguous()
    return _Linear.apply(x, weight, bias, int(act), float(slope), grad_out, step_ctx)


# ------------------------------------------------------------------------------------------------------------------------------
# Loss (reference: classification/image.py:70 — torch.nn.CrossEntropyLoss on the head's outputs)

_UNIT_GRADS = {}


def unit_grad(device) -> torch.Tensor:
    """ THE scalar 1.0 of `device`, handed to `torch.autograd.backward(loss, grad_tensors=[unit_grad(dev)])` by the captured training step: autograd then
    does not launch a fill kernel for `ones_like(loss)`, and the loss's backward recognises the tensor (by address) and skips multiplying its gradient by
    one. Never written to. """
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    t = _UNIT_GRADS.get(key)
    if t is None:
        t = _UNIT_GRADS[key] = torch.ones((), dtype=torch.float32, device=device)
    return t


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        _require_cuda(logits, target)
        m, n = logits.shape
        logits = _cast_raw(logits.contiguous(), torch.float32)
        target = target.contiguous()
        if target.dtype != torch.int64:
            raise TypeError(f'deepcv_b200: cross entropy targets must be int64 class indices, got {target.dtype}')
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        check(lib.dcv_softmax_ce(_ptr(logits), _ptr(target), _ptr(loss), _ptr(dlogits), m, n, _stream()), 'softmax_ce')
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        dlogits, = ctx.saved_tensors
        if g.data_ptr() == unit_grad(g.device).data_ptr():
            return dlogits, None   # upstream gradient is the constant 1
        g = _cast_raw(g.detach().contiguous(), torch.float32)
        out = torch.empty_like(dlogits)
        check(lib.dcv_scale_by_device_scalar(_ptr(dlogits), _ptr(g), _ptr(out), dlogits.numel(), _stream()), 'scale_by_device_scalar')
        return out, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """ `torch.nn.CrossEntropyLoss()(logits, target)` (mean reduction, class-index targets). """
    return _CrossEntropy.apply(logits, target)


# ------------------------------------------------------------------------------------------------------------------------------
# Preprocess / augmentation (reference: meta/data/preprocess.py:44-57; conf/base/parameters.yml:197-210; meta/data/augmentation.py:39-44)

def preprocess_u8(img: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, flip: Optional[torch.Tensor] = None, crop_yx: Optional[torch.Tensor] = None,
                  pad: int = 0, out_hw: Optional[Tuple[int, int]] = None, dtype: torch.dtype = torch.float32, channels_last: bool = True,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ uint8 `N x H x W x C` (device) -> normalised float `N x C x h x w` (NHWC memory unless `channels_last=False`).
    crop (zero padding `pad`, per-sample `crop_yx[n] = (top, left)`) -> horizontal flip (`flip[n] != 0`) -> (u8/255 - mean)/std. """
    _require_cuda(img, mean, std, flip, crop_yx)
    if img.dtype != torch.uint8 or img.dim() != 4 or not img.is_contiguous():
        raise TypeError(f'deepcv_b200: preprocess_u8 expects a contiguous uint8 N x H x W x C tensor, got {img.dtype} {tuple(img.shape)}')
    n, h, w, c = img.shape
    oh, ow = (h, w) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    if flip is not None and (flip.dtype != torch.uint8 or flip.numel() != n):
        raise TypeError('deepcv_b200: `flip` must be a uint8 tensor with one entry per image')
    if crop_yx is not None and (crop_yx.dtype != torch.int32 or tuple(crop_yx.shape) != (n, 2) or not crop_yx.is_contiguous()):
        raise TypeError('deepcv_b200: `crop_yx` must be a contiguous int32 N x 2 tensor')
    if out is None:
        out = empty_nhwc(n, c, oh, ow, dtype, img.device) if channels_last else torch.empty((n, c, oh, ow), dtype=dtype, device=img.device)
    check(lib.dcv_preprocess_u8(_ptr(img), _ptr(out), n, h, w, c, oh, ow, int(pad), _ptr(mean), _ptr(std), _ptr(flip), _ptr(crop_yx),
                                _dt(out), c, 0 if channels_last else 1, _stream()), 'preprocess_u8')
    return out
