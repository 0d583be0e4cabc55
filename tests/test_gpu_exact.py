""" Bit-exact parity of every tensor-core convolution kernel against `torch.nn.functional.conv2d` (and its autograd) on the CPU
(run on the B200: `pytest -m gpu`).

Operands are small integers (x, dy in {-2..2}, w in {-1, 0, 1}, integer bias): every product and every partial sum is an integer far below
2^24, so fp32 accumulation is exact in ANY summation order — tcgen05 accumulators in tensor memory, split-pixel atomics of the weight
gradient, the CPU's blocked loops all produce the same fp32 number — and the bf16 output is the round-to-nearest-even of that number on both
sides. A descriptor, swizzle, tap-offset, tile-overhang or zero-fill error of any size therefore shows up as a mismatch, not as "1 % inside a
2e-2 tolerance". Shapes: the real layers of the 224 x 224 ResNet-style spec (conf/base/resnet_style.yml, BASELINE.json configs[3]) at batch 2-3,
plus ragged maps that exercise tile overhang in every kernel variant (per-tap, resident-weight, halo, gather; per-tap and channel-block weight
gradient). Reference call site: `torch.nn.Conv2d(**submodule_params)` built at /root/reference/src/deepcv/meta/submodule_creators.py:251. """
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from deepcv_b200._lib import check, lib
    check(lib.dcv_device_check(), 'device_check')
    return torch.device('cuda', 0)


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ints(shape, lo, hi, g):
    return torch.randint(lo, hi + 1, shape, generator=g).float()


def _nhwc(t, dev, dtype=torch.bfloat16):
    """ logical NCHW tensor -> device tensor with NHWC memory """
    return t.permute(0, 2, 3, 1).contiguous().to(dev, dtype)


def _assert_equal(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    bad = (got != ref)
    assert not bool(bad.any()), f'{what}: {int(bad.sum())} of {bad.numel()} values differ (max |diff| {float((got - ref).abs().max())}, first at {tuple(bad.nonzero()[0].tolist())})'


def _check_epilogue_statistics(run, yd, y_ref, n, k, dev):
    """ BatchNorm statistics of the tcgen05 forward kernels. Flag 0: per-(image, channel) sums of the stored values (statistics kernel behind the
    convolution). DCV_STATS_CHANNEL_TOTALS (2): per-CHANNEL totals, credited to image 0 (rows of the other images stay zero) — what a BatchNorm-only
    block needs — from the statistics kernel walking the batch as one image; | DCV_STATS_IN_EPILOGUE (6): produced by the convolution epilogue. Sums of integers are exact in fp32 (< 2^24); the squares may exceed that: 1e-6 relative. Same output tensor. """
    yb = y_ref.detach().bfloat16().float()
    for flags in (0, 2, 6):
        stats = torch.full((n, k, 2), 7., device=dev)
        yd2 = torch.full_like(yd, 7.)
        run(yd2, stats, flags)
        assert torch.equal(yd2, yd), f'forward output changed with statistics flags {flags}'
        st_ = stats.cpu()
        if flags == 0:
            s1, s2, r1, r2 = st_[..., 0], st_[..., 1], yb.sum((2, 3)), (yb * yb).sum((2, 3))
        else:
            assert float(st_[1:].abs().max()) == 0. if n > 1 else True, 'channel totals must be credited to image 0 only'
            s1, s2, r1, r2 = st_[0, :, 0], st_[0, :, 1], yb.sum((0, 2, 3)), (yb.double() * yb.double()).sum((0, 2, 3)).float()
        if float(r1.abs().max()) < 2 ** 24:
            _assert_equal(s1, r1, f'sum y (flags {flags})')
        assert float((s2.double() - r2.double()).abs().max()) <= 1e-6 * float(r2.abs().max()) + 1e-3, f'sum y^2 (flags {flags})'


# n, c, h, w, k, ksize, pad — stride 1: the TMA-fed kernels
TC_EXACT = [
    (3, 64, 56, 56, 64, 3, 1),      # halo variant (resident weights, one box per tile), 4 launches per C4 step forward, 4 more as data gradient
    (2, 64, 28, 28, 128, 3, 1),     # N_TILE 128
    (3, 128, 28, 28, 128, 3, 1),
    (2, 128, 14, 14, 256, 3, 1),    # N_TILE 256
    (3, 256, 14, 14, 256, 3, 1),
    (3, 256, 7, 7, 512, 3, 1),
    (3, 512, 7, 7, 512, 3, 1),      # 7 x 7 maps: pixel tiles span several images
    (2, 64, 30, 21, 64, 3, 1),      # ragged halo tiles
    (2, 64, 9, 9, 64, 3, 1),        # resident-weight per-tap kernel, fewer pixels than one wave
    (2, 128, 40, 24, 64, 3, 1),     # two channel blocks through the halo kernel
    (5, 64, 12, 20, 64, 1, 0),      # 1 x 1 filters
    (2, 192, 14, 14, 64, 1, 0),     # channel-block weight gradient (3 blocks in one MMA)
    (2, 64, 10, 10, 64, 5, 1),      # output smaller than input (pad < (k - 1) / 2)
]


@pytest.mark.parametrize('n,c,h,w,k,ks,pad', TC_EXACT, ids=lambda v: str(v))
@pytest.mark.parametrize('act', ['none', 'relu'])
def test_tcgen05_forward_dgrad_wgrad_bit_exact(dev, n, c, h, w, k, ks, pad, act):
    from deepcv_b200._lib import ACT_NONE, ACT_RELU, ALGO_TCGEN05, DCV_BF16, ConvShape, check, lib
    g = torch.Generator().manual_seed(n * 1000 + c + k + h)
    p, q = h + 2 * pad - ks + 1, w + 2 * pad - ks + 1
    x = _ints((n, c, h, w), -2, 2, g).requires_grad_(True)
    wt = _ints((k, c, ks, ks), -1, 1, g).requires_grad_(True)
    bias = _ints((k,), -3, 3, g)
    dy = _ints((n, k, p, q), -2, 2, g)
    pre = F.conv2d(x, wt, bias, padding=pad)
    y_ref = pre.relu() if act == 'relu' else pre
    shape = ConvShape(n, h, w, c, k, ks, ks, 1, 1, pad, pad, 1, 1, p, q)
    for op in (0, 1):
        assert lib.dcv_conv2d_tc_supported(ctypes.byref(shape), DCV_BF16, op) == 1, f'op {op} not on the tcgen05 path'
    has_tc_wgrad = lib.dcv_conv2d_tc_supported(ctypes.byref(shape), DCV_BF16, 2) == 1
    assert has_tc_wgrad or ks > 3, 'the tcgen05 weight gradient serves filters up to 3 taps wide'
    st = stream()
    xd, dyd = _nhwc(x.detach(), dev), _nhwc(dy, dev)
    wd = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)          # [K][R][S][C]
    w32 = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev)
    bd = bias.to(dev)
    # ---- forward (+ bias + activation epilogue, bf16 store)
    yd = torch.full((n, p, q, k), 7., device=dev, dtype=torch.bfloat16)
    check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xd), P(wd), P(bd), P(yd), None, ACT_RELU if act == 'relu' else ACT_NONE, 0., DCV_BF16, ALGO_TCGEN05, 0, st), 'conv2d_fwd')
    _assert_equal(yd.permute(0, 3, 1, 2), y_ref.detach().bfloat16(), 'forward')
    _check_epilogue_statistics(lambda yd2, stats, flags: check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xd), P(wd), P(bd), P(yd2), P(stats), ACT_RELU if act == 'relu' else ACT_NONE, 0.,
                                                                                 DCV_BF16, ALGO_TCGEN05, flags, st), 'conv2d_fwd'), yd, y_ref, n, k, dev)
    if act == 'relu':
        return   # the backward kernels do not depend on the activation
    # ---- data and weight gradients of the linear convolution
    pre.backward(dy)
    wtd = torch.empty((c, ks, ks, k), device=dev, dtype=torch.bfloat16)
    check(lib.dcv_pack_conv_weight(P(w32), P(wtd), DCV_BF16, k, ks, ks, c, 1, st), 'pack_conv_weight')
    dxd = torch.full((n, h, w, c), 7., device=dev, dtype=torch.bfloat16)
    check(lib.dcv_conv2d_dgrad(ctypes.byref(shape), P(dyd), P(wd), P(wtd), P(dxd), DCV_BF16, ALGO_TCGEN05, st), 'conv2d_dgrad')
    _assert_equal(dxd.permute(0, 3, 1, 2), x.grad.bfloat16(), 'data gradient')
    if has_tc_wgrad:
        dwd = torch.full((k, ks, ks, c), 7., device=dev)
        check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xd), P(dyd), P(dwd), None, DCV_BF16, ALGO_TCGEN05, 0, st), 'conv2d_wgrad')
        _assert_equal(dwd.permute(0, 3, 1, 2), wt.grad, 'weight gradient (fp32)')


# n, c, h, w, k, ksize, stride, pad — the gather kernels (software im2col tile in shared memory): few input channels / strides
GATHER_EXACT = [
    (2, 3, 224, 224, 64, 7, 2, 3),   # the stem of the ImageNet-shaped net, real size
    (3, 3, 48, 40, 64, 7, 2, 3),     # ragged rows (q = 20 < 128: one partial tile per row)
    (2, 8, 20, 24, 128, 5, 1, 2),    # two output-channel atoms
    (3, 3, 8, 264, 64, 3, 1, 1),     # rows longer than two tiles
    (2, 16, 24, 24, 64, 3, 2, 1),
]


@pytest.mark.parametrize('n,c,h,w,k,ks,stride,pad', GATHER_EXACT, ids=lambda v: str(v))
def test_gather_forward_wgrad_bit_exact(dev, n, c, h, w, k, ks, stride, pad):
    from deepcv_b200._lib import ACT_NONE, DCV_BF16, ConvShape, check, lib
    g = torch.Generator().manual_seed(c * 100 + h + k)
    p, q = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
    x = _ints((n, c, h, w), -2, 2, g)
    wt = _ints((k, c, ks, ks), -1, 1, g).requires_grad_(True)
    bias = _ints((k,), -3, 3, g)
    dy = _ints((n, k, p, q), -2, 2, g)
    y_ref = F.conv2d(x, wt, bias, stride=stride, padding=pad)
    y_ref.backward(dy)
    shape = ConvShape(n, h, w, c, k, ks, ks, stride, stride, pad, pad, 1, 1, p, q)
    sc = ks * c
    kpad = (ks * ((sc + 7) // 8 * 8) + 63) // 64 * 64
    xd, dyd = _nhwc(x, dev), _nhwc(dy, dev)
    if not lib.dcv_conv2d_gather_supported(ctypes.byref(shape), P(xd), kpad, DCV_BF16):
        pytest.skip('shape not served by the gather kernels')
    st = stream()
    wd = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    w_col = torch.empty((k, kpad), device=dev, dtype=torch.bfloat16)
    check(lib.dcv_gather_pack_weight(P(wd), P(w_col), k, ks, sc, kpad, DCV_BF16, st), 'gather_pack_weight')
    yd = torch.full((n, p, q, k), 7., device=dev, dtype=torch.bfloat16)
    bd = bias.to(dev)
    check(lib.dcv_conv2d_fwd_gather(ctypes.byref(shape), P(xd), P(w_col), kpad, P(bd), P(yd), None, ACT_NONE, 0., 0, st), 'conv2d_fwd_gather')
    _assert_equal(yd.permute(0, 3, 1, 2), y_ref.detach().bfloat16(), 'gather forward')
    _check_epilogue_statistics(lambda yd2, stats, flags: check(lib.dcv_conv2d_fwd_gather(ctypes.byref(shape), P(xd), P(w_col), kpad, P(bd), P(yd2), P(stats), ACT_NONE, 0., flags, st),
                                                               'conv2d_fwd_gather'), yd, y_ref, n, k, dev)
    dw_col = torch.full((k, kpad), 7., device=dev)
    check(lib.dcv_conv2d_wgrad_gather(ctypes.byref(shape), P(xd), P(dyd), P(dw_col), kpad, 0, st), 'conv2d_wgrad_gather')
    dwd = torch.full((k, ks, ks, c), 7., device=dev)
    check(lib.dcv_gather_unpack_wgrad(P(dw_col), P(dwd), k, ks, sc, kpad, st), 'gather_unpack_wgrad')
    _assert_equal(dwd.permute(0, 3, 1, 2), wt.grad, 'gather weight gradient (fp32)')


# n, c, h, w, k, ksize, pad — stride 2, at most four input channels: the pixel-pair kernels (overlapping tensor-core tiles over pair-transposed rows)
PAIRS_EXACT = [
    (2, 3, 224, 224, 64, 7, 3),    # the stem of the ImageNet-shaped net, real size
    (3, 3, 48, 40, 64, 7, 3),      # ragged rows (q = 20 < 128: one partial tile per row)
    (2, 4, 36, 32, 128, 5, 2),     # four channels (a whole 16-byte chunk per pair), two output-channel atoms, even padding
    (2, 2, 10, 528, 64, 3, 1),     # rows longer than two tiles
    (3, 1, 16, 32, 64, 7, 3),      # one channel
    (2, 3, 30, 64, 64, 8, 4),      # eight filter rows, eight taps
]


@pytest.mark.parametrize('n,c,h,w,k,ks,pad', PAIRS_EXACT, ids=lambda v: str(v))
def test_pixel_pair_forward_wgrad_bit_exact(dev, n, c, h, w, k, ks, pad):
    from deepcv_b200._lib import ACT_NONE, DCV_BF16, ConvShape, check, lib
    g = torch.Generator().manual_seed(c * 100 + h + k)
    stride = 2
    p, q = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
    x = _ints((n, c, h, w), -2, 2, g)
    wt = _ints((k, c, ks, ks), -1, 1, g).requires_grad_(True)
    bias = _ints((k,), -3, 3, g)
    dy = _ints((n, k, p, q), -2, 2, g)
    y_ref = F.conv2d(x, wt, bias, stride=stride, padding=pad)
    y_ref.backward(dy)
    shape = ConvShape(n, h, w, c, k, ks, ks, stride, stride, pad, pad, 1, 1, p, q)
    xd, dyd = _nhwc(x, dev), _nhwc(dy, dev)
    assert lib.dcv_conv2d_pairs_supported(ctypes.byref(shape), P(xd), DCV_BF16), 'shape not served by the pixel-pair kernels'
    st = stream()
    wd = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    w_col = torch.full((k, 256), 7., device=dev, dtype=torch.bfloat16)
    check(lib.dcv_pairs_pack_weight(P(wd), P(w_col), ctypes.byref(shape), DCV_BF16, st), 'pairs_pack_weight')
    assert float(w_col.float().abs().sum()) == float(wt.detach().abs().sum())   # every weight exactly once, zeros elsewhere
    yd = torch.full((n, p, q, k), 7., device=dev, dtype=torch.bfloat16)
    bd = bias.to(dev)
    check(lib.dcv_conv2d_fwd_pairs(ctypes.byref(shape), P(xd), P(w_col), P(bd), P(yd), None, ACT_NONE, 0., 0, st), 'conv2d_fwd_pairs')
    _assert_equal(yd.permute(0, 3, 1, 2), y_ref.detach().bfloat16(), 'pixel-pair forward')
    stats = torch.full((n, k, 2), 7., device=dev)
    check(lib.dcv_conv2d_fwd_pairs(ctypes.byref(shape), P(xd), P(w_col), P(bd), P(yd), P(stats), ACT_NONE, 0., 0, st), 'conv2d_fwd_pairs(stats)')
    yf = y_ref.detach().bfloat16().double()
    assert torch.equal(stats[..., 0].double().cpu(), yf.sum((2, 3))), 'sum y'
    r2 = (yf * yf).sum((2, 3))
    assert float((stats[..., 1].double().cpu() - r2).abs().max()) <= 1e-6 * float(r2.max()) + 1e-3, 'sum y^2'
    dw_col = torch.full((k, 256), 7., device=dev)
    check(lib.dcv_conv2d_wgrad_pairs(ctypes.byref(shape), P(xd), P(dyd), P(dw_col), 0, st), 'conv2d_wgrad_pairs')
    dwd = torch.full((k, ks, ks, c), 7., device=dev)
    check(lib.dcv_pairs_unpack_wgrad(P(dw_col), P(dwd), ctypes.byref(shape), st), 'pairs_unpack_wgrad')
    _assert_equal(dwd.permute(0, 3, 1, 2), wt.grad, 'pixel-pair weight gradient (fp32)')


@pytest.mark.parametrize('n,c,h,w,k,ks,pad', [(4, 3, 32, 32, 4, 5, 2), (4, 4, 32, 32, 4, 5, 2), (3, 4, 16, 16, 16, 3, 1), (3, 16, 16, 16, 16, 3, 1), (2, 5, 13, 11, 6, 3, 1)], ids=lambda v: str(v))
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
def test_direct_kernels_bit_exact(dev, n, c, h, w, k, ks, pad, dtype):
    """ The CUDA-core kernels of the few-channel layers (default CIFAR-10 net) under the same integer-operand argument. """
    from deepcv_b200._lib import ACT_NONE, ALGO_DIRECT, DCV_BF16, DCV_F32, ConvShape, check, lib
    dt = DCV_F32 if dtype == torch.float32 else DCV_BF16
    g = torch.Generator().manual_seed(c * 7 + k)
    p, q = h + 2 * pad - ks + 1, w + 2 * pad - ks + 1
    x = _ints((n, c, h, w), -2, 2, g).requires_grad_(True)
    wt = _ints((k, c, ks, ks), -1, 1, g).requires_grad_(True)
    bias = _ints((k,), -3, 3, g)
    dy = _ints((n, k, p, q), -2, 2, g)
    y_ref = F.conv2d(x, wt, bias, padding=pad)
    y_ref.backward(dy)
    shape = ConvShape(n, h, w, c, k, ks, ks, 1, 1, pad, pad, 1, 1, p, q)
    st = stream()
    xd, dyd = _nhwc(x.detach(), dev, dtype), _nhwc(dy, dev, dtype)
    wd = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev, dtype)
    bd = bias.to(dev)
    yd = torch.full((n, p, q, k), 7., device=dev, dtype=dtype)
    stats = torch.full((n, k, 2), 7., device=dev)
    check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xd), P(wd), P(bd), P(yd), P(stats), ACT_NONE, 0., dt, ALGO_DIRECT, 0, st), 'conv2d_fwd')
    _assert_equal(yd.permute(0, 3, 1, 2), y_ref.detach().to(dtype), 'direct forward')
    if dtype == torch.float32:   # fused per-(image, channel) statistics of the stored values: integer sums, exact
        _assert_equal(stats[..., 0], y_ref.detach().sum((2, 3)), 'sum y')
        _assert_equal(stats[..., 1], (y_ref.detach() ** 2).sum((2, 3)), 'sum y^2')
    dxd = torch.full((n, h, w, c), 7., device=dev, dtype=dtype)
    check(lib.dcv_conv2d_dgrad(ctypes.byref(shape), P(dyd), P(wd), None, P(dxd), dt, ALGO_DIRECT, st), 'conv2d_dgrad')
    _assert_equal(dxd.permute(0, 3, 1, 2), x.grad.to(dtype), 'direct data gradient')
    dwd = torch.full((k, ks, ks, c), 7., device=dev)
    check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xd), P(dyd), P(dwd), None, dt, ALGO_DIRECT, 0, st), 'conv2d_wgrad')
    _assert_equal(dwd.permute(0, 3, 1, 2), wt.grad, 'direct weight gradient')


@pytest.mark.parametrize('n,c,h,w,k,ks', [(5, 3, 32, 32, 4, 5), (4, 4, 32, 32, 4, 5), (3, 4, 16, 16, 16, 3), (3, 16, 16, 16, 16, 3), (2, 16, 32, 32, 4, 3), (2, 4, 48, 32, 4, 3), (600, 4, 16, 16, 4, 3), (7, 4, 16, 32, 4, 5), (5, 16, 32, 16, 16, 3)],
                         ids=lambda v: str(v))
@pytest.mark.parametrize('act', ['none', 'relu'])
def test_few_channel_mma_kernels_bit_exact(dev, n, c, h, w, k, ks, act):
    """ The fused few-channel kernels of the default CIFAR-10 net (warp-level mma.sync implicit GEMM straight from the NHWC tile, csrc/conv_small.cu)
    with a plain input and no normalisation: forward (+ bias + activation + per-(image, channel) statistics), data gradient and weight / bias
    gradient, all bit-exact under the integer-operand argument. n = 600 > 4 x 148 CTAs: the persistent image loop. """
    from deepcv_b200._lib import ACT_NONE, ACT_RELU, DCV_BF16, ConvShape, ScNorm, check, lib
    g = torch.Generator().manual_seed(n + c * 10 + k)
    pad = ks // 2
    x = _ints((n, c, h, w), -2, 2, g).requires_grad_(True)
    wt = _ints((k, c, ks, ks), -1, 1, g).requires_grad_(True)
    bias = _ints((k,), -3, 3, g).requires_grad_(True)
    dz = _ints((n, k, h, w), -2, 2, g)
    pre = F.conv2d(x, wt, bias, padding=pad)
    y_ref = pre.relu() if act == 'relu' else pre
    y_ref.backward(dz)
    shape = ConvShape(n, h, w, c, k, ks, ks, 1, 1, pad, pad, 1, 1, h, w)
    assert lib.dcv_sc_conv_supported(ctypes.byref(shape), DCV_BF16) == 1
    st = stream()
    act_code = ACT_RELU if act == 'relu' else ACT_NONE
    xd, dzd = _nhwc(x.detach(), dev), _nhwc(dz, dev)
    wd = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    bd = bias.detach().to(dev)
    yd = torch.full((n, h, w, k), 7., device=dev, dtype=torch.bfloat16)
    # statistics only (no BatchNorm, no GroupNorm): a descriptor whose sums are produced but never used
    stats = torch.full((n, k, 2), 7., device=dev)
    nd = ScNorm(1, n, k, h * w, 0, 0, 1e-5, 0.1, None, None, None, None, None, 0, 1, 1e-5, None, None, P(stats), None, None, None)
    check(lib.dcv_sc_conv_fwd(ctypes.byref(shape), P(xd), None, 0, P(wd), P(bd), act_code, 0., P(yd), ctypes.byref(nd), st), 'sc_conv_fwd')
    _assert_equal(yd.permute(0, 3, 1, 2), y_ref.detach().bfloat16(), 'forward')
    yb = y_ref.detach().bfloat16().float()
    _assert_equal(stats[..., 0], yb.sum((2, 3)), 'sum y')
    if float(yb.abs().max()) < 64:    # squares of bf16-exact integers stay exact in fp32 sums
        _assert_equal(stats[..., 1], (yb * yb).sum((2, 3)), 'sum y^2')
    dxd = torch.full((n, h, w, c), 7., device=dev, dtype=torch.bfloat16)
    check(lib.dcv_sc_conv_dgrad(ctypes.byref(shape), P(dzd), P(yd), None, act_code, 0., P(wd), P(dxd), None, None, st), 'sc_conv_dgrad')
    _assert_equal(dxd.permute(0, 3, 1, 2), x.grad.bfloat16(), 'data gradient')
    dwd, dbd = torch.zeros((k, ks, ks, c), device=dev), torch.zeros((k,), device=dev)
    check(lib.dcv_sc_conv_wgrad(ctypes.byref(shape), P(xd), None, P(dzd), P(yd), None, act_code, 0., P(dwd), P(dbd), None, None, None, None, st), 'sc_conv_wgrad')
    _assert_equal(dwd.permute(0, 3, 1, 2), wt.grad, 'weight gradient')
    _assert_equal(dbd, bias.grad, 'bias gradient')
    # both gradients in one launch (the dy tile staged once): the layers of the default net that need a data gradient
    if lib.dcv_sc_conv_bwd_supported(ctypes.byref(shape), DCV_BF16):
        dx2 = torch.full((n, h, w, c), 7., device=dev, dtype=torch.bfloat16)
        dw2, db2 = torch.zeros((k, ks, ks, c), device=dev), torch.zeros((k,), device=dev)
        check(lib.dcv_sc_conv_bwd(ctypes.byref(shape), P(xd), None, P(dzd), P(yd), None, act_code, 0., P(wd), P(dx2), P(dw2), P(db2), None, None, None, None, st), 'sc_conv_bwd')
        _assert_equal(dx2.permute(0, 3, 1, 2), x.grad.bfloat16(), 'fused backward: data gradient')
        _assert_equal(dw2.permute(0, 3, 1, 2), wt.grad, 'fused backward: weight gradient')
        _assert_equal(db2, bias.grad, 'fused backward: bias gradient')
    else:
        assert (c, k, ks) not in ((4, 4, 5), (4, 16, 3), (16, 16, 3))
