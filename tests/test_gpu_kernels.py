""" Kernel-level parity through the C ABI (run on the B200: `pytest -m gpu`): the streaming BatchNorm / GroupNorm kernels, the row-staged im2col and
the tcgen05 weight gradient on the shapes their work decomposition treats specially — ranges that cross image boundaries, channel counts smaller
than a 16-byte vector (packed pixels), scalar fallbacks, predicated tails, 1x1 filters whose channel blocks share one MMA. Each is checked against
the same arithmetic written with stock torch ops in fp64 on the CPU (integer / index work bit-exact, floating point within the stated tolerance). """
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


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


# (n, hw, c): many images with few pixels (several images per CTA), few images with many pixels (several CTAs per image), packed 1/2/4-channel
# vectors, channel counts that force the scalar path, more vector columns than threads
NORM_SHAPES = [(512, 1024, 4), (512, 256, 16), (7, 49, 512), (3, 3136, 64), (33, 5, 3), (2, 10, 2), (4, 6, 1), (5, 77, 24), (2, 9, 4104), (300, 1, 8), (1, 40000, 8)]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('n,hw,c', NORM_SHAPES)
def test_norm_streaming_kernels(dev, n, hw, c, dtype):
    """ dcv_norm_stats / _apply_fwd / _bwd_reduce / dcv_act_norm_bwd_apply against fp64 torch on the values as stored (fp32 accumulation on the
    device: 1e-5 relative on the sums; the maps are exact up to one rounding of the output type). """
    from deepcv_b200._lib import ACT_LEAKY_RELU, DCV_BF16, DCV_F32, check, lib
    dt = DCV_F32 if dtype == torch.float32 else DCV_BF16
    g = torch.Generator().manual_seed(n * 131 + hw * 7 + c)
    y = torch.randn(n, hw, c, generator=g).to(dtype)
    dz = torch.randn(n, hw, c, generator=g).to(dtype)
    ab = torch.randn(n, c, 2, generator=g)
    pqr = torch.randn(n, c, 3, generator=g)
    yd, dzd, abd, pqrd = y.to(dev), dz.to(dev), ab.to(dev), pqr.to(dev)
    y64, dz64 = y.double(), dz.double()
    st = stream()
    # statistics
    stats = torch.full((n, c, 2), 7., device=dev)
    check(lib.dcv_norm_stats(P(yd), P(stats), n, hw, c, dt, 0, st), 'norm_stats')
    ref = torch.stack([y64.sum(1), (y64 * y64).sum(1)], -1)
    assert rel(stats, ref) <= 1e-5
    # forward apply
    z = torch.empty_like(yd)
    check(lib.dcv_norm_apply_fwd(P(yd), P(abd), P(z), n, hw, c, dt, st), 'norm_apply_fwd')
    zref = (ab[:, None, :, 0].double() * y64 + ab[:, None, :, 1].double())
    tol = 1e-6 if dtype == torch.float32 else 2.0 ** -8
    assert float(((z.double().cpu() - zref).abs() / (zref.abs() + 1.)).max()) <= tol
    # backward reduce
    s_nc = torch.full((n, c, 2), -3., device=dev)
    check(lib.dcv_norm_bwd_reduce(P(dzd), P(yd), P(s_nc), n, hw, c, dt, 0, st), 'norm_bwd_reduce')
    ref = torch.stack([dz64.sum(1), (dz64 * y64).sum(1)], -1)
    assert rel(s_nc, ref) <= 1e-5
    # backward apply + bias gradient
    dy = torch.empty_like(yd)
    dbias = torch.full((c,), 11., device=dev)
    check(lib.dcv_act_norm_bwd_apply(P(dzd), P(yd), P(pqrd), P(dy), P(dbias), ACT_LEAKY_RELU, 0.01, n, hw, c, dt, 0, st), 'act_norm_bwd_apply')
    pre = pqr[:, None, :, 0].double() * dz64 + pqr[:, None, :, 1].double() * y64 + pqr[:, None, :, 2].double()
    dyref = pre * torch.where(y64 > 0, 1.0, 0.01)
    assert float(((dy.double().cpu() - dyref).abs() / (dyref.abs() + 1.)).max()) <= tol
    assert rel(dbias, dyref.sum((0, 1))) <= 1e-4   # summed in fp32 from the unrounded values, whatever the output type
    # without parameters (P = 1, Q = R = 0) and without a bias gradient
    check(lib.dcv_act_norm_bwd_apply(P(dzd), P(yd), None, P(dy), None, ACT_LEAKY_RELU, 0.01, n, hw, c, dt, 0, st), 'act_norm_bwd_apply')
    assert float(((dy.double().cpu() - dz64 * torch.where(y64 > 0, 1.0, 0.01)).abs()).max()) <= (1e-6 if dtype == torch.float32 else 2.0 ** -7 * float(dz64.abs().max()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('n,c,h,w,r,s,stride,pad,dil', [(2, 3, 224, 224, 7, 7, 2, 3, 1), (3, 3, 33, 29, 7, 7, 2, 3, 1), (2, 5, 17, 19, 3, 3, 1, 1, 1), (2, 16, 24, 24, 3, 3, 2, 1, 1),
                                                        (2, 8, 20, 20, 3, 3, 1, 2, 2), (1, 1, 9, 40, 5, 3, 3, 0, 1), (2, 4, 12, 12, 1, 1, 1, 0, 1)])
def test_im2col_bit_exact(dev, n, c, h, w, r, s, stride, pad, dil, dtype):
    """ dcv_im2col (row-staged kernel) against torch unfold: pure data movement, bit-exact, zero in the padding and in the columns beyond R*S*C. """
    from deepcv_b200._lib import DCV_BF16, DCV_F32, ConvShape, check, lib
    dt = DCV_F32 if dtype == torch.float32 else DCV_BF16
    p = (h + 2 * pad - dil * (r - 1) - 1) // stride + 1
    q = (w + 2 * pad - dil * (s - 1) - 1) // stride + 1
    g = torch.Generator().manual_seed(h * w + c)
    x = torch.randn(n, c, h, w, generator=g).to(dtype)
    rsc = r * s * c
    kpad = (rsc + 63) // 64 * 64
    col = torch.full((n, p, q, kpad), 5., dtype=dtype, device=dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(dev)
    shape = ConvShape(n, h, w, c, 64, r, s, stride, stride, pad, pad, dil, dil, p, q)
    check(lib.dcv_im2col(ctypes.byref(shape), P(x_nhwc), P(col), kpad, dt, stream()), 'im2col')
    unf = F.unfold(x.float(), (r, s), dilation=dil, padding=pad, stride=stride)           # [n][c*r*s][p*q], rows ordered (c, r, s)
    ref = unf.view(n, c, r, s, p, q).permute(0, 4, 5, 2, 3, 1).reshape(n, p, q, rsc)       # -> (r, s, c) fastest c
    got = col.float().cpu()
    assert torch.equal(got[..., :rsc], ref)
    assert float(got[..., rsc:].abs().max()) == 0. if kpad > rsc else True


@pytest.mark.parametrize('n,c,hw,k', [(4, 192, 14, 64), (3, 128, 9, 128), (2, 256, 12, 64), (2, 64, 16, 128), (5, 320, 7, 64)])
def test_tcgen05_wgrad_pointwise_channel_blocks(dev, n, c, hw, k):
    """ 1x1 filters: the weight gradient groups 3 or 2 consecutive 64-channel blocks of x into one N = 192 / 128 MMA (the im2col GEMM of the stem has
    c = kpad = 192). bf16 operands, fp32 accumulation: 2e-2 of max|dw| against fp64 on the same bf16 values (measured ~1e-3). """
    from deepcv_b200._lib import ALGO_TCGEN05, DCV_BF16, ConvShape, check, lib
    g = torch.Generator().manual_seed(c + k)
    x = torch.randn(n, hw, hw, c, generator=g).bfloat16()
    dy = torch.randn(n, hw, hw, k, generator=g).bfloat16()
    shape = ConvShape(n, hw, hw, c, k, 1, 1, 1, 1, 0, 0, 1, 1, hw, hw)
    if not lib.dcv_conv2d_tc_supported(ctypes.byref(shape), DCV_BF16, 2):
        pytest.skip('shape not on the tcgen05 weight-gradient path')
    dw = torch.full((k, 1, 1, c), 9., device=dev)
    xd, dyd = x.to(dev), dy.to(dev)   # keep the device copies alive across the asynchronous launch
    check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xd), P(dyd), P(dw), None, DCV_BF16, ALGO_TCGEN05, 0, stream()), 'conv2d_wgrad')
    torch.cuda.synchronize()
    ref = dy.double().reshape(-1, k).t() @ x.double().reshape(-1, c)
    assert rel(dw.reshape(k, c), ref) <= 2e-2
    assert rel(dw.reshape(k, c), ref) <= 5e-3, 'fp32 accumulation of bf16 products should be far inside the bf16 tolerance'


@pytest.mark.parametrize('momentum', [0.07359778246238029, None])
@pytest.mark.parametrize('n,hw,c,groups', [(6, 49, 96, 0), (512, 16, 4, 4), (3, 100, 256, 32), (2, 9, 2048, 0)])
def test_norm_finalize_running_statistics_and_affine(dev, n, hw, c, groups, momentum):
    """ dcv_norm_fwd_finalize over several steps against torch.nn.BatchNorm2d (+ GroupNorm): running_mean / running_var / num_batches_tracked (also with
    momentum=None, the cumulative average, where every channel's update reads the counter: served by a single CTA) and the folded per-(n,c) affine
    applied to y reproduces BatchNorm -> GroupNorm of the reference modules. """
    from deepcv_b200._lib import DCV_F32, NormParams, check, lib
    torch.manual_seed(c + n)
    bn = torch.nn.BatchNorm2d(c, eps=1e-5, momentum=momentum)
    gn = torch.nn.GroupNorm(groups, c) if groups else None
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
        if gn is not None:
            gn.weight.uniform_(0.5, 1.5); gn.bias.normal_()
    rm, rv = bn.running_mean.clone().to(dev), bn.running_var.clone().to(dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    bw, bb = bn.weight.detach().to(dev), bn.bias.detach().to(dev)
    gw = gn.weight.detach().to(dev) if gn is not None else None
    gb = gn.bias.detach().to(dev) if gn is not None else None
    side = int(round(hw ** 0.5))
    assert side * side == hw
    st = stream()
    for step in range(3):
        y = torch.randn(n, c, side, side) * (1.0 + step) + 0.3 * step
        ref = bn(y)
        if gn is not None:
            ref = gn(ref)
        yd = y.permute(0, 2, 3, 1).contiguous().to(dev)                     # NHWC
        stats = torch.empty(n, c, 2, device=dev)
        check(lib.dcv_norm_stats(P(yd), P(stats), n, hw, c, DCV_F32, 0, st), 'norm_stats')
        G = groups if groups else 1
        saved = torch.empty(int(lib.dcv_norm_saved_floats(n, c, G)), device=dev)
        ab = torch.empty(n, c, 2, device=dev)
        prm = NormParams(n, c, hw, 1, 1, 1e-5, -1.0 if momentum is None else momentum, P(bw), P(bb), P(rm), P(rv), P(nbt),
                         1 if groups else 0, G, 1e-5, P(gw), P(gb))
        check(lib.dcv_norm_fwd_finalize(ctypes.byref(prm), P(stats), P(ab), P(saved), st), 'norm_fwd_finalize')
        z = torch.empty_like(yd)
        check(lib.dcv_norm_apply_fwd(P(yd), P(ab), P(z), n, hw, c, DCV_F32, st), 'norm_apply_fwd')
        torch.cuda.synchronize()
        assert int(nbt) == step + 1 == int(bn.num_batches_tracked)
        assert rel(rm, bn.running_mean) <= 1e-5 and rel(rv, bn.running_var) <= 1e-5
        assert rel(z.permute(0, 3, 1, 2), ref.detach()) <= 1e-4


def test_batched_transposed_weight_pack(dev):
    """ `dcv_pack_conv_weights_batched`: the data-gradient operands [C][R-1-r][S-1-s][K] (bf16) of several [K][R][S][C] fp32 weights living in one flat buffer,
    in one launch — bit-equal to the per-layer `dcv_pack_conv_weight` and to the permute / flip / round-to-bf16 of the logical OIHW tensor. """
    from deepcv_b200 import ops
    torch.manual_seed(0)
    shapes = [(64, 64, 3, 3), (128, 64, 3, 3), (48, 128, 1, 1), (16, 4, 5, 5), (10, 3, 7, 7), (512, 256, 3, 3), (33, 20, 3, 2)]   # (K, C, R, S)
    offs, total = [], 0
    for k, c, r, s in shapes:
        offs.append(total)
        total += (k * c * r * s + 7) // 8 * 8
    flat = torch.randn(total, device=dev)
    weights = [flat[o:o + k * c * r * s].view(k, r, s, c).permute(0, 3, 1, 2) for o, (k, c, r, s) in zip(offs, shapes)]     # logical OIHW, memory KRSC
    ctx = ops.StepContext()
    ctx.plan_transposed_weights(flat, weights, torch.bfloat16)
    assert ctx.transposed is not None and ctx.transposed['n'] == len(shapes)
    ctx.refresh_shadows()
    for w, (k, c, r, s) in zip(weights, shapes):
        got = ctx.transposed_view(w, torch.bfloat16)
        assert tuple(got.shape) == (c, r, s, k)
        ref = w.permute(1, 2, 3, 0).flip(1, 2).contiguous().bfloat16()          # [C][R-1-r][S-1-s][K]
        assert torch.equal(got, ref), (k, c, r, s)
    assert ctx.transposed_view(torch.randn(4, 4, 3, 3, device=dev), torch.bfloat16) is None


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('n,h,w,channels,pools,needs', [
    (512, 8, 8, (16, 4), (1, 2), (True, True)),             # the default net's dense link (conf/base/parameters.yml:86)
    (3, 5, 7, (3, 5, 2), (1, 2, 1), (True, False, True)),   # odd channel counts: scalar granules; a source that needs no gradient
    (2, 6, 4, (8, 8, 16, 8, 24, 8, 8, 8), (1, 1, 2, 1, 1, 2, 1, 1), (True,) * 8),
    (4, 3, 3, (6, 2), (2, 1), (True, True)),
])
def test_dense_link_concat_in_one_launch(dev, n, h, w, channels, pools, needs, dtype):
    """ ops.link_concat_rescaled == F.interpolate(bilinear, align_corners=False) of the twice-larger tensors + torch.cat, bit for bit each way. """
    from deepcv_b200 import ops
    torch.manual_seed(sum(channels) + h)
    srcs = [torch.randn(n, c, h * p, w * p, device=dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(r)
            for c, p, r in zip(channels, pools, needs)]
    if pools[0] != 1:
        assert ops.link_concat_rescaled(srcs) is None   # the previous output defines the spatial shape: it is never rescaled
        return
    out = ops.link_concat_rescaled(srcs)
    refs = [s.detach().float().requires_grad_(r) for s, r in zip(srcs, needs)]
    parts = [F.interpolate(r, size=(h, w), mode='bilinear', align_corners=False).to(dtype).float() if p == 2 else r for r, p in zip(refs, pools)]
    want = torch.cat(parts, dim=1)
    assert out.shape == want.shape and torch.equal(out.float(), want)
    g = torch.randn_like(want).to(dtype)
    out.backward(g.contiguous(memory_format=torch.channels_last))
    off = 0
    for s, c, p, r in zip(srcs, channels, pools, needs):
        gs = g[:, off:off + c].float()
        off += c
        if not r:
            assert s.grad is None
            continue
        expect = (gs * 0.25).to(dtype).float().repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) if p == 2 else gs
        assert torch.equal(s.grad.float(), expect)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('n,h,w,c', [(3, 8, 12, 64), (2, 6, 4, 24), (5, 2, 2, 3), (2, 56, 56, 64)])
def test_norm_apply_fused_with_residual_sum_and_pooling(dev, n, h, w, c, dtype):
    """ dcv_norm_apply_add_fwd = A*y + B + other and dcv_norm_apply_pool_fwd = A*avgpool2x2(y) + B against the same arithmetic in fp64 (one rounding to the
    storage type at the end: 2^-8 relative for bf16, 1e-6 for fp32), and `ops.apply_pending` backward = the gradient w.r.t. the normalised tensor. """
    from deepcv_b200 import ops
    from deepcv_b200._lib import DCV_BF16, DCV_F32, check, lib
    torch.manual_seed(n * 100 + c)
    dt = DCV_BF16 if dtype == torch.bfloat16 else DCV_F32
    y = torch.randn(n, h, w, c, device=dev).to(dtype)
    other = torch.randn(n, h, w, c, device=dev).to(dtype)
    ab = torch.randn(n, c, 2, device=dev)
    A, B = ab[..., 0].double()[:, None, None, :], ab[..., 1].double()[:, None, None, :]
    tol = 2 ** -8 if dtype == torch.bfloat16 else 1e-6
    out = torch.full_like(y, 7.)
    check(lib.dcv_norm_apply_add_fwd(P(y), P(ab), P(other), P(out), n, h * w, c, dt, stream()), 'norm_apply_add_fwd')
    want = A * y.double() + B + other.double()
    assert float(((out.double() - want).abs() / (want.abs() + 1.)).max()) <= tol
    outp = torch.full((n, h // 2, w // 2, c), 7., device=dev, dtype=dtype)
    check(lib.dcv_norm_apply_pool_fwd(P(y), P(ab), P(outp), n, h, w, c, dt, stream()), 'norm_apply_pool_fwd')
    wantp = F.avg_pool2d((A * y.double() + B).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert float(((outp.double() - wantp).abs() / (wantp.abs() + 1.)).max()) <= tol
    # autograd contract of a pending normalisation: the consumer returns dz (gradient w.r.t. A*y + B) in the slot of y
    yl = y.permute(0, 3, 1, 2).requires_grad_(True)
    ol = other.permute(0, 3, 1, 2).requires_grad_(True)
    z = ops.apply_pending(ops.PendingNorm(yl, ab), other=ol)
    g = torch.randn_like(z)
    z.backward(g)
    assert torch.equal(yl.grad, g) and torch.equal(ol.grad, g)
    yl.grad = None
    zp = ops.apply_pending(ops.PendingNorm(yl, ab), pool=True)
    gp = torch.randn_like(zp)
    zp.backward(gp)
    assert torch.equal(yl.grad.float(), (gp.float() * 0.25).to(dtype).float().repeat_interleave(2, dim=2).repeat_interleave(2, dim=3))


@pytest.mark.parametrize('totals', [0, 2], ids=['per-image', 'channel-totals'])
@pytest.mark.parametrize('n,h,w,c', [(3, 8, 12, 64), (2, 6, 4, 24), (4, 112, 112, 64)])
def test_norm_backward_reads_pooled_gradient_in_place(dev, n, h, w, c, totals):
    """ dcv_norm_bwd_reduce_pooled / dcv_act_norm_bwd_apply_pooled == dcv_avgpool2d_bwd followed by the plain passes: dy bit for bit, the sums to fp32
    summation-order noise (atomics). """
    from deepcv_b200._lib import ACT_LEAKY_RELU, DCV_BF16, check, lib
    torch.manual_seed(h * 10 + c)
    st = stream()
    gp = torch.randn(n, h // 2, w // 2, c, device=dev).bfloat16()
    y = torch.randn(n, h, w, c, device=dev).bfloat16()
    pqr = torch.randn(n, c, 3, device=dev)
    assert lib.dcv_norm_bwd_pooled_supported(n, h, w, c, DCV_BF16)
    dz = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
    check(lib.dcv_avgpool2d_bwd(P(gp), P(dz), n, h, w, c, 2, 2, 2, 2, DCV_BF16, st), 'avgpool2d_bwd')
    s_ref, s_new = torch.full((n, c, 2), 7., device=dev), torch.full((n, c, 2), 7., device=dev)
    check(lib.dcv_norm_bwd_reduce(P(dz), P(y), P(s_ref), n, h * w, c, DCV_BF16, totals, st), 'norm_bwd_reduce')
    check(lib.dcv_norm_bwd_reduce_pooled(P(gp), P(y), P(s_new), n, h, w, c, DCV_BF16, totals, st), 'norm_bwd_reduce_pooled')
    assert float((s_ref - s_new).abs().max()) <= 2e-5 * float(s_ref.abs().max()) + 1e-6
    dy_ref, dy_new = torch.full_like(y, 7.), torch.full_like(y, 7.)
    db_ref, db_new = torch.full((c,), 7., device=dev), torch.full((c,), 7., device=dev)
    check(lib.dcv_act_norm_bwd_apply(P(dz), P(y), P(pqr), P(dy_ref), P(db_ref), ACT_LEAKY_RELU, 0.01, n, h * w, c, DCV_BF16, 0, st), 'act_norm_bwd_apply')
    check(lib.dcv_act_norm_bwd_apply_pooled(P(gp), P(y), P(pqr), P(dy_new), P(db_new), ACT_LEAKY_RELU, 0.01, n, h, w, c, DCV_BF16, 0, st), 'act_norm_bwd_apply_pooled')
    assert torch.equal(dy_ref, dy_new)
    assert float((db_ref - db_new).abs().max()) <= 2e-5 * float(db_ref.abs().max()) + 1e-6
