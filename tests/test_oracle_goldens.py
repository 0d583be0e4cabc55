""" The CPU oracle against the committed golden fixtures (tests/golden/, made by oracle/make_goldens.py): reference pure functions lifted
from /root/reference, real torchvision ToTensor/Normalize/crop/hflip outputs, and the oracle's own frozen network outputs. """
import copy
import hashlib
import json

import numpy as np
import pytest
import torch

from oracle import deepcv_oracle as O


@pytest.fixture(scope='module')
def ref_fns(golden_dir):
    return json.loads((golden_dir / 'ref_pure_functions.json').read_text())


def test_padding_from_kernel_matches_reference(ref_fns):
    from deepcv_b200.meta.nn import get_padding_from_kernel
    for k, ref in ref_fns['get_padding_from_kernel'].items():
        ks = json.loads(k)
        # the reference returns the first dim's padding (an int torch broadcasts); for these square kernels that equals the per-dim list
        assert O.get_padding_from_kernel(ks) == [ref] * len(ks)
        assert get_padding_from_kernel(ks) == [ref] * len(ks)
        conv_a, conv_b = torch.nn.Conv2d(1, 1, ks, padding=ref), torch.nn.Conv2d(1, 1, ks, padding=get_padding_from_kernel(ks))
        assert conv_a.padding == conv_b.padding


def test_get_by_identifier_matches_reference(ref_fns):
    from deepcv_b200.utils import get_by_identifier
    for ident, qual in ref_fns['get_by_identifier'].items():
        obj = get_by_identifier(ident)
        assert f'{obj.__module__}.{obj.__qualname__}' == qual
    for ident, err in ref_fns['get_by_identifier_errors'].items():
        with pytest.raises((ValueError, RuntimeError, ImportError, AttributeError)) as e:
            get_by_identifier(ident)
        if err in ('ValueError', 'RuntimeError'):
            assert type(e.value).__name__ == err


def test_to_hyperparameters_matches_reference(ref_fns):
    from deepcv_b200.meta.hyperparams import to_hyperparameters
    for case in ref_fns['to_hyperparameters']:
        defaults = {k: (... if v == '...' else v) for k, v in case['defaults'].items()}
        hp, missing = to_hyperparameters(case['hp'], defaults, raise_if_missing=False, drop_keys_not_in_defaults=case['drop'])
        assert dict(hp) == case['result'] and missing == case['missing']
    with pytest.raises(ValueError, match='Missing mandatory'):
        to_hyperparameters({'b': 1}, {'a': ...})


def test_preprocess_oracle_matches_torchvision(golden_dir):
    gold = torch.load(golden_dir / 'torchvision_preprocess.pt')
    for name in ('cifar', 'imagenet'):
        g = gold[name]
        plain = O.preprocess_u8(g['images'], g['mean'], g['std'])
        assert torch.equal(plain, g['plain']), name          # same fp32 op order as ToTensor -> Normalize: bit-identical
        aug = O.preprocess_u8(g['images'], g['mean'], g['std'], flip=g['flip'], crop_yx=g['crop'], pad=g['pad'])
        assert torch.equal(aug, g['augmented']), name        # RandomCrop(padding, fill=0) + hflip geometry, zero pad normalises to (0-mean)/std
    ramp = gold['ramp']
    assert torch.equal(O.preprocess_u8(ramp['images'], [0.491, 0.482, 0.447], [0.247, 0.243, 0.261]), ramp['plain'])


def test_index_maps_golden(golden_dir):
    gold = json.loads((golden_dir / 'index_maps.json').read_text())
    assert len(gold) == 2 * 9 * 9
    for key, digest in gold.items():
        flip, top, left = map(int, key.split(','))
        m = O.preprocess_index_map(32, 32, 32, 32, 4, flip, top, left)
        assert hashlib.sha256(np.ascontiguousarray(m.astype('<i4')).tobytes()).hexdigest() == digest
    # the index map is what preprocess_u8 actually selects
    img = torch.arange(32 * 32, dtype=torch.int32).remainder(251).to(torch.uint8).view(1, 32, 32, 1)
    for flip, top, left in ((0, 0, 0), (1, 8, 3), (1, 4, 4), (0, 7, 8)):
        out = O.preprocess_u8(img, [0.], [1.], flip=torch.tensor([flip], dtype=torch.uint8), crop_yx=torch.tensor([[top, left]], dtype=torch.int32), pad=4)
        m = O.preprocess_index_map(32, 32, 32, 32, 4, flip, top, left)
        expect = np.where(m[..., 0] >= 0, img[0, :, :, 0].numpy()[np.clip(m[..., 0], 0, 31), np.clip(m[..., 1], 0, 31)], 0)
        assert np.array_equal((out[0, 0] * 255).round().numpy().astype(np.uint8), expect.astype(np.uint8))


def test_default_net_shape_table_and_capacity(default_hp):
    """ per-submodule output shapes and the 17 010 parameter count. """
    model = O.OracleDeepcvModule((3, 32, 32), default_hp)
    backbone = model._submodules['_submodule_0']
    assert backbone._features_shapes[1:] == [(4, 32, 32), (4, 32, 32), (4, 32, 32), (4, 16, 16), (16, 16, 16), (16, 16, 16), (16, 8, 8), (20, 8, 8)]
    assert model._features_shapes[1:] == [(20, 8, 8), (1280,), (10,)]
    assert sum(p.numel() for p in model.parameters()) == 17010
    block = backbone._submodules['_submodule_0']
    assert [type(m).__name__ for m in block] == ['Conv2d', 'ReLU', 'BatchNorm2d', 'GroupNorm']   # conv -> act -> BN -> GN (post-activation norm)
    assert block[2].momentum == 0.07359778246238029 and block[3].num_groups == 4


def test_oracle_reproduces_frozen_network_outputs(golden_dir, default_hp):
    gold = torch.load(golden_dir / 'oracle_default_net.pt')
    model = O.OracleDeepcvModule(gold['input_shape'], default_hp)
    model.load_state_dict(gold['state'])
    loss, logits = O.train_step(model, gold['x'], gold['y'])
    assert abs(loss - gold['loss']) <= 1e-5 * abs(gold['loss'])
    torch.testing.assert_close(logits, gold['logits'], rtol=1e-4, atol=1e-6)
    for n, p in model.named_parameters():
        ref = gold['grads'][n]
        assert (p.grad - ref).abs().max() <= 1e-4 * ref.abs().max() + 1e-9, n
    for n, v in model.state_dict().items():
        torch.testing.assert_close(v, gold['state_after'][n], rtol=1e-5, atol=1e-7)


def test_bilinear_halving_is_2x2_average():
    """ F.interpolate(bilinear, align_corners=False) to exactly half size == avg_pool2d(2). (NOT true for 4x.) """
    x = torch.randn(2, 3, 16, 16)
    a = torch.nn.functional.interpolate(x, size=(8, 8), mode='bilinear', align_corners=False)
    assert (a - torch.nn.functional.avg_pool2d(x, 2)).abs().max() < 1e-6
    b = torch.nn.functional.interpolate(x, size=(4, 4), mode='bilinear', align_corners=False)
    assert (b - torch.nn.functional.avg_pool2d(x, 4)).abs().max() > 1e-3


def test_link_semantics():
    link = O.Link('sum', allow_scaling=False)
    link.referenced_submodules = ['a']
    x, a = torch.randn(1, 2, 4, 4), torch.randn(1, 2, 4, 4)
    assert torch.equal(link(x, {'a': a}), x + a)
    cat = O.Link('concat', allow_scaling=True)
    cat.referenced_submodules = ['a']
    big = torch.randn(1, 3, 8, 8)
    out = cat(x, {'a': big})
    assert out.shape == (1, 5, 4, 4) and torch.equal(out[:, :2], x)
    with pytest.raises(RuntimeError):
        bad = O.Link('concat', allow_scaling=False)
        bad.referenced_submodules = ['a']
        bad(x, {'a': big})
