import copy, sys, torch
sys.path.insert(0, '.')
from deepcv_b200.yaml_config import find_model_spec, load_parameters
from deepcv_b200.meta.base_module import DeepcvModule
from deepcv_b200.meta.ignite_training import CrossEntropyLoss
from oracle.deepcv_oracle import OracleDeepcvModule, train_step
dev = torch.device('cuda')
hp = dict(find_model_spec(load_parameters('conf/base/parameters.yml'), 'image_classifier'))
hp['architecture'] = copy.deepcopy(hp['architecture']); hp['architecture'][-1]['fully_connected']['out_features'] = 10
def rel(a, b): return float((a.float().cpu() - b.float()).abs().max() / max(float(b.abs().max()), 1e-12))
for batch in (8, 128, 512):
    torch.manual_seed(1)
    oracle = OracleDeepcvModule((3, 32, 32), hp)
    model = DeepcvModule((3, 32, 32), hp); model.load_state_dict(oracle.state_dict()); model = model.to(dev)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, 3, 32, 32, generator=g).bfloat16().float(); y = torch.randint(0, 10, (batch,), generator=g)
    loss_ref, logits_ref = train_step(oracle, x, y)
    o64 = copy.deepcopy(oracle).double(); train_step(o64, x.double(), y)
    for dtype in (torch.float32, torch.bfloat16):
        model.zero_grad(); model.train()
        logits = model(x.to(dev, dtype)); loss = CrossEntropyLoss()(logits, y.to(dev)); loss.backward()
        errs = {n.replace('_child_modules.', '').replace('_submodule_', 's'): rel(p.grad, dict(oracle.named_parameters())[n].grad) for n, p in model.named_parameters()}
        print(batch, dtype, 'logits', f'{rel(logits, logits_ref):.2e}', 'loss', f'{abs(float(loss)-loss_ref)/loss_ref:.2e}')
        print('   ', {k: f'{v:.1e}' for k, v in errs.items()})
