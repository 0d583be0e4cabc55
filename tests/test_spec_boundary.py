""" Host-side boundary (no GPU): YAML grammar, `!py!` tags, creator registry and calling convention, error behaviour, state_dict
interchange with the oracle, shape inference on the `meta` device — mirrors the reference's parse-time semantics (nn_spec.py, submodule_creators.py). """
import copy
from collections import OrderedDict
from pathlib import Path

import pytest
import torch

from deepcv_b200.meta import nn as dnn
from deepcv_b200.meta import nn_spec, submodule_creators
from deepcv_b200.meta.base_module import DeepcvModule
from deepcv_b200.yaml_config import find_model_spec, load_parameters, loads_parameters
from oracle import deepcv_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_yaml_tags_and_floats():
    p = load_parameters(ROOT / 'conf' / 'base' / 'parameters.yml')
    hp = find_model_spec(p, 'image_classifier')
    assert hp['act_fn'] is torch.nn.LeakyReLU and hp['batch_norm']['eps'] == 1e-05 and isinstance(hp['batch_norm']['eps'], float)
    assert hp['architecture'][1] is torch.nn.Flatten
    assert p['train_image_classifier']['optimizer_opts']['lr'] == 1e-3
    import torchvision
    norm = p['cifar10_preprocessing']['transforms'][1]
    assert list(norm.keys())[0] is torchvision.transforms.Normalize


@pytest.mark.skipif(not Path('/root/reference/conf/base/parameters.yml').exists(), reason='reference tree only exists in the build container')
def test_reference_parameters_yml_parses_unchanged():
    p = load_parameters('/root/reference/conf/base/parameters.yml')
    hp = dict(find_model_spec(p, 'image_classifier'))
    assert hp['spectral_norm'] is not None       # the reference sets it (and then raises on it): benchmark configs null it
    hp['spectral_norm'] = None
    hp['architecture'] = copy.deepcopy(hp['architecture'])
    hp['architecture'][-1]['fully_connected']['out_features'] = 10
    model = DeepcvModule((3, 32, 32), hp)
    assert sum(p.numel() for p in model.parameters()) == 17010
    # the shipped conf/base/parameters.yml IS the reference's file
    assert (ROOT / 'conf' / 'base' / 'parameters.yml').read_bytes() == Path('/root/reference/conf/base/parameters.yml').read_bytes()


def test_default_net_structure_matches_oracle(default_hp):
    model = DeepcvModule((3, 32, 32), default_hp)
    oracle = O.OracleDeepcvModule((3, 32, 32), default_hp)
    backbone = model._submodules['_submodule_0']
    assert list(backbone._features_shapes) == list(oracle._submodules['_submodule_0']._features_shapes)
    assert list(model._features_shapes) == list(oracle._features_shapes)
    sd_m, sd_o = model.state_dict(), oracle.state_dict()
    assert list(sd_m.keys()) == list(sd_o.keys())
    assert all(sd_m[k].shape == sd_o[k].shape for k in sd_m)
    model.load_state_dict(sd_o)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd_o[k])
    conv_w = backbone._submodules['_submodule_1'][0].weight
    assert conv_w.permute(0, 2, 3, 1).is_contiguous()      # physically [K][R][S][C], logically OIHW
    assert not model.is_sequential_nn(recursive=True) and model.is_sequential_nn() and not backbone.is_sequential_nn()
    assert isinstance(model._submodules['_submodule_1'], dnn.Flatten)


def test_resnet_style_spec_builds():
    hp = find_model_spec(load_parameters(ROOT / 'conf' / 'base' / 'resnet_style.yml'), 'resnet_style_classifier')
    model = DeepcvModule((3, 224, 224), hp)
    assert model._features_shapes[-1] == (1000,) and model._features_shapes[1] == (768, 1, 1)
    assert sum(p.numel() for p in model.parameters()) == 10153192
    bb = model._submodules['_submodule_0']
    assert bb._features_shapes[1] == (64, 112, 112) and bb._submodule_references['s1b1'] == ['s1in']


def test_initialisation_follows_reference(default_hp):
    torch.manual_seed(0)
    model = DeepcvModule((3, 32, 32), default_hp)
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            assert float(m.bias.detach().abs().max()) == 0.
        if isinstance(m, torch.nn.BatchNorm2d):
            assert 0. <= float(m.weight.detach().min()) and float(m.weight.detach().max()) <= 1. and float(m.bias.detach().abs().max()) == 0.
    # Xavier-normal with the gain of the OUTER act_fn (LeakyReLU), also for the nested backbone's ReLU convolutions (base_module.py:237,264)
    conv = model._submodules['_submodule_0']._submodules['_submodule_5'][0]
    gain = torch.nn.init.calculate_gain('leaky_relu')
    expect_std = gain * (2. / (16 * 9 + 16 * 9)) ** 0.5
    assert abs(float(conv.weight.detach().std()) - expect_std) < 0.15 * expect_std


def test_creator_calling_convention():
    seen = {}

    def my_creator(submodule_params, input_shape, prev_shapes, act_fn, width: int = 3):
        seen.update(submodule_params=submodule_params, input_shape=input_shape, prev_shapes=list(prev_shapes), act_fn=act_fn, width=width)
        return torch.nn.Identity()

    hp = {'act_fn': torch.nn.ReLU, 'width': 5, 'architecture': [{'my_op': {'foo': 1, 'width': 7}}, {'my_op': ['named', {'bar': 2}]}, 'torch.nn.Identity']}
    model = DeepcvModule((3, 8, 8), hp, additional_submodule_creators={'my_op': my_creator})
    assert list(model._submodules.keys()) == ['_submodule_0', 'named', '_submodule_2']
    assert seen['submodule_params'] == {'bar': 2} and seen['width'] == 5 and seen['input_shape'] == (3, 8, 8) and seen['act_fn'] is torch.nn.ReLU
    # user creators override the built-in ones (nn_spec.py:76-77)
    hp2 = {'act_fn': torch.nn.ReLU, 'architecture': [{'conv2d': {'kernel_size': [3, 3], 'out_channels': 4}}]}
    m2 = DeepcvModule((3, 8, 8), hp2, additional_submodule_creators={'conv2d': lambda submodule_params: torch.nn.Identity()})
    assert isinstance(m2._submodules['_submodule_0'], torch.nn.Identity)


def test_spec_errors():
    base = {'act_fn': torch.nn.ReLU}
    with pytest.raises(ValueError, match='Missing mandatory'):
        DeepcvModule((3, 8, 8), {'architecture': []})
    with pytest.raises(ValueError, match='duplicate'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'avg_pooling': ['a', {'kernel_size': [2, 2]}]}, {'avg_pooling': ['a', {'kernel_size': [2, 2]}]}]})
    with pytest.raises(RuntimeError, match='Could not locate'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'no_such_creator': {}}]})
    with pytest.raises(ValueError, match='Invalid sub-module reference'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'residual_link': {'_from': 'nowhere'}}]})
    with pytest.raises(ValueError, match='Missing "_from"'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'residual_link': {}}]})
    with pytest.raises(RuntimeError, match='must either be a parameters Dict'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'avg_pooling': 3}]})
    with pytest.raises(RuntimeError, match='allow_scaling'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'avg_pooling': ['p', {'kernel_size': [1, 1]}]}, {'avg_pooling': {'kernel_size': [2, 2]}}, {'dense_link': {'_from': 'p'}}]})
    with pytest.raises(NotImplementedError, match='hot path'):
        DeepcvModule((3, 8, 8), {**base, 'architecture': [{'_nas_layer_choice': {'_candidates': []}}]})
    # dropout and the pre-activation order are built (reference nn.py:535-541, 553): same child modules in the reference's order
    m = DeepcvModule((4, 8, 8), {**base, 'dropout_prob': 0.5, 'preactivation': True, 'batch_norm': {'affine': True, 'eps': 1e-5, 'momentum': 0.1},
                                 'architecture': [{'conv2d': {'kernel_size': [3, 3], 'out_channels': 6}}]})
    block = m._submodules['_submodule_0']
    assert [type(c).__name__ for c in block] == ['Dropout', 'BatchNorm2d', type(base['act_fn']()).__name__, 'Conv2d'] and block[1].num_features == 4
    m = DeepcvModule((4, 8, 8), {**base, 'dropout_prob': 0.5, 'batch_norm': {'affine': True, 'eps': 1e-5, 'momentum': 0.1},
                                 'architecture': [{'conv2d': {'kernel_size': [3, 3], 'out_channels': 6}}]})
    block = m._submodules['_submodule_0']
    assert [type(c).__name__ for c in block] == ['Dropout', 'Conv2d', type(base['act_fn']()).__name__, 'BatchNorm2d'] and block[3].num_features == 6
    with pytest.raises(NotImplementedError, match='Dropout'):
        from deepcv_b200.meta.nn import FusedLayer
        FusedLayer(torch.nn.Dropout2d(0.5), torch.nn.Conv2d(3, 4, 3))


def test_registry_decorator():
    reg = {}

    @submodule_creators.submodule_creator_dec('thing', submodule_creators=reg, allowed_subm_params_keys={'a'}, required_subm_params_keys={'b'})
    def thing(submodule_params):
        thing._check_submodule_params(submodule_params)
        return torch.nn.Identity()
    assert reg['thing'] is thing
    thing({'a': 1, 'b': 2})
    with pytest.raises(ValueError, match='not allowed'):
        thing({'b': 1, 'zzz': 2})
    with pytest.raises(ValueError, match='Missing'):
        thing({'a': 1})
    with pytest.raises(AssertionError):
        submodule_creators.submodule_creator_dec('thing', submodule_creators=reg)
    assert {'conv2d', 'fully_connected', 'linear', 'average_pooling', 'avg_pooling', 'residual_link', 'dense_link', 'reduce', '_new_branch_from_tensor'} <= set(submodule_creators.BASIC_SUBMODULE_CREATORS)


def test_layer_order_and_norm_factory():
    blk = dnn.layer(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.LeakyReLU, input_shape=(3, 16, 16), batch_norm={'eps': 1e-5, 'momentum': 0.07}, group_norm={'num_groups': 4})
    assert [type(m).__name__ for m in blk] == ['Conv2d', 'LeakyReLU', 'BatchNorm2d', 'GroupNorm'] and isinstance(blk, torch.nn.Sequential)
    assert blk[2].num_features == 8 and blk[3].num_channels == 8
    with pytest.raises(ValueError, match='no `weight`'):
        dnn.layer(torch.nn.ReLU(), torch.nn.ReLU)
    with pytest.raises(ValueError, match='mutiple times'):
        dnn.normalization_techniques_impl(['batch_norm', 'batch_norm'], [{}, {}], input_shape=(4, 2, 2))
    assert dnn.is_conv(torch.nn.Conv2d) and dnn.is_conv(torch.nn.Conv3d(1, 1, 1)) and not dnn.is_conv(torch.nn.Linear)   # reference nn.py:731-740
    assert dnn.get_gain_name(torch.nn.LeakyReLU) == 'leaky_relu' and dnn.get_gain_name(torch.nn.GELU) == 'relu'


def test_preprocess_recipe_api(golden_dir):
    from deepcv_b200.meta.data import preprocess as P
    params = load_parameters(ROOT / 'conf' / 'base' / 'parameters.yml')
    gold = torch.load(golden_dir / 'torchvision_preprocess.pt')['cifar']

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return gold['images'].shape[0]

        def __getitem__(self, i):
            from PIL import Image
            return Image.fromarray(gold['images'][i].numpy()), int(i)

    recipe = dict(params['cifar10_preprocessing'])
    recipe['split_dataset'] = {'validset_ratio': None, 'testset_ratio': 0.34}
    out = P.preprocess(recipe, DS(), None)
    assert set(out) == {'trainset', 'testset'} and isinstance(out['trainset'], P.PreprocessedDataset)
    x, y = out['trainset'][0]
    assert torch.equal(x, gold['plain'][y])                  # the reference recipe: torchvision ToTensor + Normalize, per sample
    fused = dict(load_parameters(ROOT / 'conf' / 'base' / 'b200.yml')['cifar10_fused_preprocessing'])
    fused['split_dataset'] = {'validset_ratio': None, 'testset_ratio': 0.34}
    out = P.preprocess(fused, DS(), None)
    x, y = out['trainset'][0]
    assert x.dtype == torch.uint8 and x.shape == (32, 32, 3) and torch.equal(x, gold['images'][y])   # per sample the fused transform keeps uint8 HWC
    tf = out['trainset']._img_transform.transforms[0]
    assert isinstance(tf, P.FusedPreprocess) and tf.pad == 4 and tf.flip
    with pytest.raises(RuntimeError, match='CUDA'):
        tf(gold['images'])
    # runtime-computed arguments (mean/std absent from the YAML) go through TRANSFORM_ARGS_PROCESSORS
    stats = P._parse_transforms_specification([{P.FusedPreprocess: {}}], trainset=DS()).transforms[0]
    ref_mean = torch.stack([gold['images'][i].float().div(255).permute(2, 0, 1).mean(dim=(1, 2)) for i in range(6)]).mean(0)
    assert torch.allclose(stats.mean, ref_mean, atol=1e-6)
    with pytest.raises(RuntimeError, match='already registered'):
        P.register_transform_processor(P.FusedPreprocess, ['mean'])(lambda trainset, to_process: {})


def test_piecewise_linear_and_engine():
    from deepcv_b200.meta.ignite_training import Engine, Events, PiecewiseLinear
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.)
    sched = PiecewiseLinear(opt, 'lr', [[0, 0.0], [4, 1.0], [10, 0.0]])
    seen = []
    eng = Engine(lambda e, b: seen.append((e.state.iteration, opt.param_groups[0]['lr'])))
    eng.add_event_handler(Events.ITERATION_STARTED, sched)
    state = eng.run(list(range(5)), max_epochs=2)
    assert state.iteration == 10 and state.epoch == 2
    assert [round(v, 4) for _, v in seen[:6]] == [0.0, 0.25, 0.5, 0.75, 1.0, round(1 - 1 / 6, 4)]


def test_accumulator_arena_counts_and_parameter_shadow_views():
    """ Host logic of the two per-step consolidations of the captured training step (no kernel involved): the accumulator arena sizes itself in
    counting mode and hands out plain tensors outside a step; a convolution weight that lives in a flat parameter buffer maps to the same element
    range of the buffer's low-precision shadow, in the same [K][R][S][C] order. """
    import torch
    from deepcv_b200 import ops
    arena = ops.AccumulatorArena()
    t = arena.alloc((3, 5, 2), torch.device('cpu'))
    assert t.shape == (3, 5, 2) and t.dtype == torch.float32 and arena.need == 0
    arena.measure()
    arena.alloc((3, 5, 2), torch.device('cpu'))
    arena.alloc((7,), torch.device('cpu'))
    assert arena.need == 256 + 256                       # 120 and 28 bytes, each rounded up to 256
    assert arena.end_measure(torch.device('cpu')) == 512 and arena.buf.numel() >= 512 and not arena.counting
    flat = torch.arange(64, dtype=torch.float32)
    shadow = flat.to(torch.bfloat16)
    k, c, r, s = 2, 3, 2, 2
    w = flat[8:8 + k * c * r * s].view(k, r, s, c).permute(0, 3, 1, 2)    # logically OIHW, physically KRSC, inside the flat buffer
    ctx = ops.StepContext(arena)
    assert ctx.prezeroed == 0                            # only between begin_step() and end_step()
    ctx.shadows = [(flat, shadow)]
    v = ctx.shadow_view(w, torch.bfloat16)
    assert v is not None and v.shape == w.shape and v.data_ptr() == shadow.data_ptr() + 8 * 2
    assert torch.equal(v.float(), w)
    assert ctx.shadow_view(torch.zeros(k, c, r, s).permute(0, 1, 2, 3), torch.bfloat16) is None      # not in the buffer
    assert ctx.shadow_view(w, torch.float16) is None                                                  # no shadow of that dtype
    # per-step state is per object: a second context (another model / stream / capturing thread) sees nothing of the first
    assert ops.StepContext().shadow_view(w, torch.bfloat16) is None and not hasattr(ops, '_CURRENT_ARENA') and not hasattr(ops, '_PARAM_SHADOWS')


def test_bench_clock_rules():
    """ bench.py timing rules (host logic only): which clock samples reject a measurement, and that only samples read inside the timed window count. """
    import importlib.util
    import sys
    from pathlib import Path
    spec = importlib.util.spec_from_file_location('dcv_bench', Path(__file__).resolve().parent.parent / 'bench.py')
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    assert not bench.needs_remeasure(dict(sm_mhz=1965., sm_max_mhz=1965., reasons=[]))
    assert not bench.needs_remeasure(dict(sm_mhz=1500., sm_max_mhz=1965., reasons=['sw_power_cap']))      # kept and noted
    assert bench.needs_remeasure(dict(sm_mhz=1965., sm_max_mhz=1965., reasons=['hw_thermal_slowdown']))
    assert bench.needs_remeasure(dict(sm_mhz=900., sm_max_mhz=1965., reasons=[]))                          # a leftover clock lock
    assert not bench.needs_remeasure(dict(sm_mhz=None, sm_max_mhz=None, reasons=[]))                       # no nvidia-smi
    cs = bench.ClockSampler(0)
    idle = ['1965', '1965', '400', 'Not Active', 'Not Active', 'Not Active', 'Not Active']
    busy = ['1200', '1965', '900', 'Not Active', 'Not Active', 'Not Active', 'Active']
    cs.rows = [(10.0, idle), (10.5, busy), (11.0, idle)]
    inside = cs.summary(10.4, 10.6)
    assert inside['samples'] == 1 and inside['sm_mhz'] == 1200. and inside['reasons'] == ['sw_power_cap']
    assert cs.summary(20., 21.)['samples'] == 1                                                             # nothing inside: the nearest sample
