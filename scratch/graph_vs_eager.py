""" Per-iteration losses of train() with the graph-replayed step vs the eager step (debug aid for test_train_entry_point_graph_replay_matches_eager). """
import copy, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import test_gpu_training as T
from deepcv_b200.classification import image
from deepcv_b200.meta import ignite_training as IT
from deepcv_b200.yaml_config import benchmark_model_spec

default_hp = benchmark_model_spec(ROOT / 'conf' / 'base' / 'parameters.yml', 'image_classifier', out_features=10)
datasets = T._datasets()
torch.manual_seed(3)
model_a = image.create_model(datasets, copy.deepcopy(default_hp))
state0 = copy.deepcopy(model_a.state_dict())
tf = IT._find_fused_preprocess(datasets['trainset'])
orig = IT.Engine.run
def run_with(graph):
    m = image.create_model(datasets, copy.deepcopy(default_hp)); m.load_state_dict(state0)
    tf.generator.manual_seed(434546)
    losses = []
    def run(self, data, max_epochs=1, epoch_length=None):
        self.add_event_handler(IT.Events.ITERATION_COMPLETED, lambda e: losses.append(e.state.output['main_loss']))
        return orig(self, data, max_epochs, epoch_length)
    IT.Engine.run = run
    try:
        image.train(datasets, m, T._hp(cuda_graph=graph))
    finally:
        IT.Engine.run = orig
    return losses, m
la, ma = run_with(None)
lb, mb = run_with(False)
lc, mc = run_with(False)
for i, (a, b, c) in enumerate(zip(la, lb, lc)):
    print(f'iter {i}: graph {a:.6f} eager {b:.6f} eager2 {c:.6f} diff {a-b:+.2e} eager-eager {b-c:+.2e}')
for (n, a), (_, b), (_, c) in zip(ma.state_dict().items(), mb.state_dict().items(), mc.state_dict().items()):
    if a.dtype.is_floating_point:
        print(f'{n:70s} graph-eager {float((a-b).abs().max()):.2e} eager-eager {float((b-c).abs().max()):.2e}  max|w| {float(b.abs().max()):.2e}')
