""" The public training path on the device (run on the B200: `pytest -m gpu`): `classification.image.create_model` / `train` ->
`ignite_training.train` with the fused preprocess recipe, the device-resident input pipeline (`DeviceDataLoader` + `dcv_gather_rows`), the
CUDA-graph replayed `process_function`, the evaluation pass (`GraphedEvalStep` + `dcv_classification_metrics`) and the loss kernel's handling
of ignored / invalid targets. Reference: /root/reference/src/deepcv/classification/image.py:40-80, meta/ignite_training.py:178-307. """
import copy
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from deepcv_b200._lib import check, lib
    check(lib.dcv_device_check(), 'device_check')
    return torch.device('cuda', 0)


class _Images(torch.utils.data.Dataset):
    classes = [str(i) for i in range(10)]

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randint(0, 256, (n, 32, 32, 3), generator=g, dtype=torch.uint8)
        self.y = torch.randint(0, 10, (n,), generator=g)
        # a learnable signal: the label sets the mean brightness of the first channel
        self.x[..., 0] = (self.x[..., 0].float() * 0.25 + self.y.view(-1, 1, 1).float() * 19).clamp(0, 255).to(torch.uint8)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i].numpy(), int(self.y[i])


def _datasets():
    from deepcv_b200.meta.data import preprocess as P
    from deepcv_b200.yaml_config import load_parameters
    recipe = dict(load_parameters(ROOT / 'conf' / 'base' / 'b200.yml')['cifar10_fused_preprocessing'])
    recipe['split_dataset'] = {'validset_ratio': None, 'testset_ratio': None}
    return P.preprocess(recipe, _Images(167, 1), _Images(70, 2))


def _hp(**over):
    from deepcv_b200.yaml_config import load_parameters
    hp = dict(load_parameters(ROOT / 'conf' / 'base' / 'parameters.yml')['train_image_classifier'])
    hp.update(epochs=2, batch_size=32, seed=563454, backend_conf={'device_or_id': 'cuda:0'})
    hp.update(over)
    return hp


def test_train_entry_point_graph_replay_matches_eager(dev, default_hp):
    """ `train(datasets, model, hp)` twice from the same initial weights: CUDA-graph replayed step (the default) and eager step. Same batches (the
    sampler is seeded), same augmentation draws (the capture rewinds the generator), same kernels: parameters agree up to atomics ordering. """
    from deepcv_b200.classification import image
    datasets = _datasets()
    assert set(datasets) == {'trainset', 'testset'}
    model_hp = copy.deepcopy({k: v for k, v in default_hp.items()})
    del model_hp['architecture'][-1]['fully_connected']['out_features']      # deduced from `trainset.classes` (reference :44-50)
    torch.manual_seed(3)
    model_a = image.create_model(datasets, model_hp)
    assert model_a._input_shape == (3, 32, 32) and model_a._submodules['_submodule_2'][0].out_features == 10
    state0 = copy.deepcopy(model_a.state_dict())
    model_b = image.create_model(datasets, model_hp)
    model_b.load_state_dict(state0)
    from deepcv_b200.meta.ignite_training import _find_fused_preprocess
    recipe_transform = _find_fused_preprocess(datasets['trainset'])
    recipe_transform.generator.manual_seed(434546)          # both runs draw the same flips / crop offsets (one transform object serves both)
    metrics_a, state_a, _ = image.train(datasets, model_a, _hp())
    recipe_transform.generator.manual_seed(434546)
    metrics_b, state_b, _ = image.train(datasets, model_b, _hp(cuda_graph=False))
    assert state_a.iteration == state_b.iteration == 2 * (167 // 32) and state_a.epoch == 2
    assert set(metrics_a) == {'valid_loss', 'valid_accuracy', 'valid_samples'} and metrics_a['valid_samples'] == 70
    # Convolution / fully connected weights (BatchNorm affine under a one-channel-per-group GroupNorm and the convolution biases in front of BatchNorm
    # have analytically zero gradients: Adam turns their rounding noise — which depends on the atomics' order — into +-lr steps on both sides). Early
    # Adam steps are ~lr * sign(g): a component whose gradient is near zero may differ by a few lr (1e-3 at the peak of the schedule) after 10 steps.
    for (n, a), (_, b) in zip(model_a.state_dict().items(), model_b.state_dict().items()):
        if a.dtype.is_floating_point and a.dim() >= 2:
            assert float((a - b).abs().max()) <= 5e-4, n       # measured 1e-5 .. 1.5e-4 (eager vs eager: 1e-8 .. 8e-5)
            assert float((a - b).abs().mean()) <= 5e-5, n
    assert abs(metrics_a['valid_loss'] - metrics_b['valid_loss']) <= 1e-3 * abs(metrics_b['valid_loss'])
    assert metrics_a['valid_accuracy'] == pytest.approx(metrics_b['valid_accuracy'], abs=2 / 70)
    assert state_a.output['main_loss'] == pytest.approx(state_b.output['main_loss'], rel=1e-5)   # second epoch, after a validation pass: same draws
    # the evaluation metrics are those of stock torch on the model's own eval-mode logits
    from deepcv_b200.meta.data.preprocess import FusedPreprocess
    test = datasets['testset']
    xs = torch.stack([torch.as_tensor(test[i][0]) for i in range(len(test))]).to(dev)
    ys = torch.tensor([test[i][1] for i in range(len(test))])
    pre = FusedPreprocess(mean=[0.491, 0.482, 0.447], std=[0.247, 0.243, 0.261], pad=4, flip=True).to(dev).eval()
    model_a.eval()
    with torch.no_grad():
        logits = model_a(pre(xs)).float().cpu()
    assert metrics_a['valid_loss'] == pytest.approx(float(F.cross_entropy(logits, ys)), rel=1e-4)
    assert metrics_a['valid_accuracy'] == pytest.approx(float((logits.argmax(1) == ys).float().mean()), abs=1e-6)


def test_training_reduces_the_loss(dev, default_hp):
    from deepcv_b200.classification import image
    datasets = _datasets()
    torch.manual_seed(5)
    model = image.create_model(datasets, default_hp)
    hp = _hp(epochs=12, scheduler=None)
    hp['optimizer_opts'] = dict(hp['optimizer_opts'], lr=1e-2)
    from deepcv_b200.meta import ignite_training as T
    first = {}
    orig = T.Engine.run

    def run(self, data, max_epochs=1, epoch_length=None):
        self.add_event_handler(T.Events.ITERATION_COMPLETED, lambda e: first.setdefault('loss', e.state.output['main_loss']))
        return orig(self, data, max_epochs, epoch_length)
    T.Engine.run = run
    try:
        metrics, state, _ = image.train(datasets, model, hp)
    finally:
        T.Engine.run = orig
    assert state.output['main_loss'] < first['loss'] - 0.05, (first, state.output)
    assert metrics['valid_accuracy'] > 0.15      # chance is 0.10; the label is encoded in the first channel's brightness


def test_device_data_loader_sampling_matches_torch(dev):
    """ Batches are rows of the dataset at exactly the indices `torch.utils.data` would draw: seeded `randperm` when shuffling,
    `DistributedSampler` sharding (wrap-around padding, rank stride) under data parallelism; `drop_last` as `DataLoader`. """
    from deepcv_b200.meta.ignite_training import DeviceDataLoader
    ds = _Images(75, 9)
    rows = [(torch.as_tensor(ds[i][0]), ds[i][1]) for i in range(len(ds))]

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return len(rows)

        def __getitem__(self, i):
            return rows[i]
    for world, rank, shuffle, drop_last in [(1, 0, False, False), (1, 0, True, True), (2, 1, True, True), (4, 3, False, False)]:
        dl = DeviceDataLoader(DS(), 16, dev, shuffle=shuffle, drop_last=drop_last, seed=11, rank=rank, world_size=world)
        dl.set_epoch(2)
        if world > 1:
            sampler = torch.utils.data.distributed.DistributedSampler(DS(), num_replicas=world, rank=rank, shuffle=shuffle, seed=11)
            sampler.set_epoch(2)
            expect = list(iter(sampler))
        elif shuffle:
            expect = torch.randperm(75, generator=torch.Generator().manual_seed(11 + 2)).tolist()
        else:
            expect = list(range(75))
        got_x, got_y = [], []
        for x, y in dl:
            assert x.dtype == torch.uint8 and x.is_cuda and y.dtype == torch.int64
            got_x.append(x.cpu().clone()), got_y.append(y.cpu().clone())
        n_batches = len(expect) // 16 if drop_last else (len(expect) + 15) // 16
        assert len(got_x) == n_batches == len(dl)
        expect = expect[: n_batches * 16] if drop_last else expect
        assert torch.equal(torch.cat(got_x), torch.stack([rows[i][0] for i in expect]))
        assert torch.cat(got_y).tolist() == [rows[i][1] for i in expect]


def test_cross_entropy_ignored_and_invalid_targets(dev):
    from deepcv_b200 import ops
    torch.manual_seed(0)
    logits = torch.randn(9, 7, requires_grad=True)
    target = torch.tensor([0, 6, -100, 3, -100, 1, 2, 5, 4])
    ref = F.cross_entropy(logits, target)
    ref.backward()
    ld = logits.detach().to(dev).requires_grad_(True)
    loss = ops.cross_entropy(ld, target.to(dev))
    loss.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert float((ld.grad.cpu() - logits.grad).abs().max()) <= 1e-6
    bad = target.clone()
    bad[1] = 7     # out of range: NaN, not an out-of-bounds read
    assert torch.isnan(ops.cross_entropy(ld.detach(), bad.to(dev)))


def test_prefetched_batches_order_values_and_overlap_safety(dev):
    """ `dataloader_prefetch_batches` (reference meta/data/datasets.py:76-115): batches arrive on the device, in order, with the right values, while the
    consumer keeps the compute stream busy (the double buffers are only rewritten after their consumers were enqueued); a ragged last batch, a
    single-tensor loader and the CPU / unpinned cases the reference returns unchanged. """
    from deepcv_b200.meta.data.datasets import PrefetchedBatches, dataloader_prefetch_batches
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randint(0, 256, (64, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory(), torch.randint(0, 10, (64,), generator=g).pin_memory()) for _ in range(7)]
    batches.append((torch.randint(0, 256, (5, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory(), torch.randint(0, 10, (5,), generator=g).pin_memory()))
    loader = dataloader_prefetch_batches(batches, dev)
    assert isinstance(loader, PrefetchedBatches) and len(loader) == 8
    busy = torch.randn(2048, 2048, device=dev)
    sums, labels = [], []
    for x, y in loader:
        assert x.is_cuda and y.is_cuda and x.dtype == torch.uint8
        for _ in range(3):
            busy = (busy @ busy).clamp_(-1, 1)          # the consumer's step: queued work that reads the batch only afterwards
        sums.append(x.sum(dtype=torch.int64) + busy[0, 0].long() * 0)
        labels.append(y.clone())
    torch.cuda.synchronize()
    assert [int(s) for s in sums] == [int(b[0].sum(dtype=torch.int64)) for b in batches]
    bad = [i for i, (l, b) in enumerate(zip(labels, batches)) if not torch.equal(l.cpu(), b[1])]
    assert not bad, (bad, [(labels[i][:6].tolist(), batches[i][1][:6].tolist()) for i in bad])
    singles = [torch.full((4, 3), float(i)).pin_memory() for i in range(3)]
    assert [float(t[0, 0]) for t in dataloader_prefetch_batches(singles, dev)] == [0., 1., 2.]
    assert dataloader_prefetch_batches(batches, 'cpu') is batches and dataloader_prefetch_batches(batches, None) is batches
    assert list(dataloader_prefetch_batches([], dev)) == []
