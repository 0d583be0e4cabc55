import copy, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import _small_resnet_hp, _oracle_runs, _run_model
from deepcv_b200.meta.base_module import DeepcvModule
from deepcv_b200.meta.ignite_training import CrossEntropyLoss
from oracle.deepcv_oracle import OracleDeepcvModule
dev = torch.device('cuda')
hp = _small_resnet_hp()
torch.manual_seed(11)
init = OracleDeepcvModule((3, 96, 96), hp)
model = DeepcvModule((3, 96, 96), hp); model.load_state_dict(init.state_dict()); model = model.to(dev)
g = torch.Generator().manual_seed(5)
x = torch.randn(6, 3, 96, 96, generator=g); y = torch.randint(0, 17, (6,), generator=g)
runs = _oracle_runs(hp, (3, 96, 96), init.state_dict(), x, y, torch.float32)
loss, logits = _run_model(model, x.to(dev), y.to(dev), CrossEntropyLoss())
def rel(a, b): return float((a.detach().double().cpu() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))
print('logits ours-64', rel(logits, runs['ref64']['logits']), 'o32-64', rel(runs['ref32']['logits'], runs['ref64']['logits']))
for n, p in model.named_parameters():
    print(n.replace('_child_modules.', '').replace('_submodule_', 's'), f"ours {rel(p.grad, runs['ref64']['grads'][n]):.2e}  oracle32 {rel(runs['ref32']['grads'][n], runs['ref64']['grads'][n]):.2e}")
