""" Both block orders of the reference's `layer()` (/root/reference/src/deepcv/meta/nn.py:519-554) and its Dropout, on the device, against the CPU
oracle (run on the B200: `pytest -m gpu`):

  post-activation  `(?Dropout) - op - act - (?norms)`          (nn.py:553, default)
  pre-activation   `(?Dropout) - (?norms) - act - op`          (nn.py:553, `preactivation: true`)
  Dropout(p)       only when `dropout_prob` not in (None, 0)   (nn.py:535-541) -> torch.nn.Dropout

Dropout parity follows SURVEY.md "Hard parts": the device draws the mask (Philox, csrc/elementwise.cu), exports it, and the oracle's `torch.nn.Dropout`
modules are replaced by modules applying exactly that mask — forward values, data gradient and every parameter gradient must then agree to the fp32
gate (1e-4, fp64 oracle as arbiter). """
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from deepcv_b200._lib import check, lib
    check(lib.dcv_device_check(), 'device_check')
    return torch.device('cuda', 0)


def _spec(preactivation: bool, dropout_prob: float, norms: bool = True):
    """ Three conv blocks (5x5 then 3x3, the second with a stride) + pooling + head, LeakyReLU, BatchNorm + GroupNorm: every layer in the given order. """
    hp = {'act_fn': torch.nn.LeakyReLU, 'dropout_prob': dropout_prob, 'preactivation': preactivation,
          'batch_norm': {'affine': True, 'eps': 1e-5, 'momentum': 0.1} if norms else None,
          'group_norm': {'num_groups': 2, 'eps': 1e-5, 'affine': True} if norms else None,
          'architecture': [{'conv2d': {'kernel_size': [5, 5], 'out_channels': 8, 'padding': 2}},
                           {'conv2d': {'kernel_size': [3, 3], 'out_channels': 12, 'padding': 1, 'stride': 2}},
                           {'conv2d': {'kernel_size': [3, 3], 'out_channels': 12, 'padding': 1}},
                           {'avg_pooling': {'kernel_size': [2, 2], 'stride': [2, 2]}},
                           'torch.nn.Flatten',
                           {'fully_connected': {'out_features': 7, 'act_fn': torch.nn.Sigmoid, 'batch_norm': None, 'group_norm': None, 'dropout_prob': 0., 'preactivation': False}}]}
    return hp


class _FixedMaskDropout(torch.nn.Module):
    """ `torch.nn.Dropout(p)` with the mask given: x * mask / (1 - p). """

    def __init__(self, p, masks):
        super().__init__()
        self.p, self.masks = p, masks

    def forward(self, x):
        if not self.training:
            return x
        mask = self.masks.pop(0).to(x.dtype)
        assert mask.shape == x.shape, (mask.shape, x.shape)
        return x * mask / (1. - self.p)


def _swap_dropout(module: torch.nn.Module, masks):
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.Dropout):
            setattr(module, name, _FixedMaskDropout(child.p, masks) if child.p != 0. else torch.nn.Identity())
        else:
            _swap_dropout(child, masks)


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize('preactivation', [False, True], ids=['postact', 'preact'])
@pytest.mark.parametrize('dropout_prob', [0., 0.3], ids=['nodrop', 'drop'])
@pytest.mark.parametrize('norms', [True, False], ids=['norms', 'plain'])
def test_block_orders_and_dropout_against_oracle(dev, preactivation, dropout_prob, norms):
    from deepcv_b200 import ops
    from deepcv_b200.meta.base_module import DeepcvModule
    from deepcv_b200.meta.nn import FusedLayer
    from oracle import deepcv_oracle as O
    hp = _spec(preactivation, dropout_prob, norms)
    torch.manual_seed(11)
    oracle = O.OracleDeepcvModule((4, 16, 16), copy.deepcopy(hp))
    model = DeepcvModule((4, 16, 16), copy.deepcopy(hp))
    assert [n for n, _ in model.named_parameters()] == [n for n, _ in oracle.named_parameters()]      # same modules, same order (state_dict interchange)
    model.load_state_dict(oracle.state_dict())
    model = model.to(dev).train()
    oracle.train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 4, 16, 16, generator=g)
    t = torch.randn(6, 7, generator=g)
    # ---- device: record the masks the Dropout kernels draw
    layers = [m for m in model.modules() if isinstance(m, FusedLayer)]
    records = []
    for layer_ in layers:
        if layer_._drop is not None and layer_._drop.p != 0.:
            layer_._drop_state = ops.DropoutState(dev, seed=1234 + len(records))
            layer_._drop_state.record = []
            records.append(layer_._drop_state.record)
    xd = x.to(dev).requires_grad_(True)
    out = model(xd)
    (out * t.to(dev)).sum().backward()
    masks = [rec[0].cpu() for rec in records]
    assert len(masks) == (3 if dropout_prob else 0)
    for m in masks:     # Bernoulli(1 - p): keep fraction within 5 sigma
        n = m.numel()
        assert abs(float(m.float().mean()) - (1. - dropout_prob)) < 5. * (dropout_prob * (1. - dropout_prob) / n) ** 0.5, float(m.float().mean())
    # ---- oracle (fp32 and fp64) with the same masks
    results = {}
    for name, dtype in (('f32', torch.float32), ('f64', torch.float64)):
        ref = copy.deepcopy(oracle).to(dtype)
        _swap_dropout(ref, [m.clone() for m in masks])
        xr = x.detach().clone().to(dtype).requires_grad_(True)
        o = ref(xr)
        (o * t.to(dtype)).sum().backward()
        results[name] = (o.detach(), xr.grad, {n: p.grad for n, p in ref.named_parameters()}, {n: b for n, b in ref.named_buffers()}, ref)
    o32, dx32, g32, b32, ref32 = results['f32']
    o64, dx64, g64, _, _ = results['f64']

    def gate(a, r32, r64, what):
        allowed = FP32_TOL * float(r64.abs().max()) + 8. * float((r32.double() - r64).abs().max()) + 1e-30
        err = float((a.detach().double().cpu() - r64).abs().max())
        assert err <= allowed, f'{what}: |err| {err:.3e} > {allowed:.3e} (rel {_rel(a, r64):.2e}; oracle fp32 vs fp64 {_rel(r32, r64):.2e})'
    gate(out, o32, o64, 'output')
    gate(xd.grad, dx32, dx64, 'input gradient')
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        gate(p.grad, g32[n], g64[n], f'grad {n}')
    for n, b in model.named_buffers():      # running statistics after one step
        if b.dtype.is_floating_point:
            assert float((b.cpu() - b32[n]).abs().max()) <= 1e-5 + 1e-5 * float(b32[n].abs().max()), n
        else:
            assert int(b) == int(b32[n]), n
    # ---- eval mode: Dropout is the identity, BatchNorm uses running statistics
    model.eval(), ref32.eval()      # ref32: the fp32 oracle after the same training-mode forward (same running statistics)
    with torch.no_grad():
        assert _rel(model(x.to(dev)), ref32(x)) <= FP32_TOL


def test_dropout_fresh_mask_per_call_and_graph_replay(dev):
    """ The call counter lives on the device: two calls draw different masks, the backward pass regenerates its forward's mask, and a CUDA-graph
    replay draws a new mask every time. """
    from deepcv_b200 import ops
    state = ops.DropoutState(dev, seed=99)
    state.record = []
    x = torch.ones(4, 8, 16, 16, device=dev, requires_grad=True)
    y1 = ops.dropout(x, 0.5, state)
    y2 = ops.dropout(x, 0.5, state)
    m1, m2 = state.record
    assert not torch.equal(m1, m2) and int(state.counter) == 2
    assert torch.equal(y1, m1.float() * 2.) and torch.equal(y2, m2.float() * 2.)
    y1.sum().backward()
    assert torch.equal(x.grad, m1.float() * 2.)
    # bf16 + ragged count (tail elements), in place semantics of the values
    xb = torch.randn(3, 5, 7, device=dev).bfloat16()
    state.record = []
    yb = ops.dropout(xb, 0.25, state)
    mb = state.record[0]
    assert torch.equal(yb, (xb.float() * mb.float() / 0.75).bfloat16())
    state.record = None
    static_x = torch.ones(1024, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.dropout(static_x, 0.5, state)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_y = ops.dropout(static_x, 0.5, state)
    outs = []
    for _ in range(3):
        graph.replay()
        outs.append(static_y.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    with pytest.raises(ValueError):
        ops.dropout(static_x, 1.0, state)


def _siamese_spec():
    """ Parallel branches (reference meta/nn.py:130-194, submodule_creators.py:175,203-224,254,300): a dense link with `reduction: none` turns two
    feature maps into a LIST of two tensors; the convolution block and the pooling that follow are applied to each of them (one shared layer:
    siamese branches); a parallel residual link adds the i-th tensor of a referenced list to the i-th branch; `reduce` concatenates the branches. """
    return {'act_fn': torch.nn.ReLU, 'dropout_prob': 0., 'batch_norm': {'affine': True, 'eps': 1e-5, 'momentum': 0.1},
            'architecture': [{'conv2d': ['a', {'kernel_size': [3, 3], 'out_channels': 8, 'padding': 1}]},
                             {'conv2d': ['b', {'kernel_size': [3, 3], 'out_channels': 8, 'padding': 1}]},
                             {'dense_link': ['pair', {'_from': 'a', 'reduction': 'none'}]},          # [b, a]: two parallel tensors
                             {'conv2d': {'kernel_size': [3, 3], 'out_channels': 8, 'padding': 1}},   # shared by both branches
                             {'residual_link': {'_from': 'pair'}},                                   # i-th branch + i-th tensor of `pair`
                             {'avg_pooling': {'kernel_size': [2, 2], 'stride': [2, 2]}},
                             {'reduce': {'fn': 'concat'}},
                             'torch.nn.Flatten',
                             {'fully_connected': {'out_features': 5, 'act_fn': torch.nn.Sigmoid, 'batch_norm': None}}]}


def test_parallel_branches_against_oracle(dev):
    from deepcv_b200.meta.base_module import DeepcvModule
    from oracle import deepcv_oracle as O
    torch.manual_seed(21)
    oracle = O.OracleDeepcvModule((4, 16, 16), _siamese_spec())
    model = DeepcvModule((4, 16, 16), _siamese_spec())
    assert model._features_shapes[3] == [(8, 16, 16), (8, 16, 16)] and model._features_shapes[4] == [(8, 16, 16), (8, 16, 16)]   # lists of shapes while branches are parallel
    assert model._features_shapes[7] == (16, 8, 8) and model._features_shapes[-1] == (5,)
    assert [n for n, _ in model.named_parameters()] == [n for n, _ in oracle.named_parameters()]
    model.load_state_dict(oracle.state_dict())
    model = model.to(dev).train()
    oracle.train()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(5, 4, 16, 16, generator=g)
    t = torch.randn(5, 5, generator=g)
    xd = x.to(dev).requires_grad_(True)
    out = model(xd)
    (out * t.to(dev)).sum().backward()
    o64m = copy.deepcopy(oracle).double()
    xr64 = x.double().requires_grad_(True)
    o64 = o64m(xr64)
    (o64 * t.double()).sum().backward()
    xr = x.clone().requires_grad_(True)
    o32 = oracle(xr)
    (o32 * t).sum().backward()

    def gate(a, r32, r64, what):
        allowed = FP32_TOL * float(r64.abs().max()) + 8. * float((r32.double() - r64).abs().max()) + 1e-30
        err = float((a.detach().double().cpu() - r64).abs().max())
        assert err <= allowed, f'{what}: |err| {err:.3e} > {allowed:.3e}'
    gate(out, o32.detach(), o64.detach(), 'output')
    gate(xd.grad, xr.grad, xr64.grad, 'input gradient')
    g64 = dict(o64m.named_parameters())
    for (n, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
        gate(p.grad, q.grad, g64[n].grad, f'grad {n}')   # the shared layer's gradients are the SUM over both branches
    # tensor-count checks of the call convention
    bad = _siamese_spec()
    bad['architecture'][4] = {'residual_link': {'_from': 'a'}}   # one referenced tensor for two parallel inputs
    with pytest.raises(ValueError, match='in_tensors_count_similar_to_refs'):
        DeepcvModule((4, 16, 16), bad)
