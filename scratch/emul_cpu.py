import copy, sys, torch
sys.path.insert(0, '.')
from deepcv_b200.yaml_config import find_model_spec, load_parameters
from oracle.deepcv_oracle import OracleDeepcvModule, train_step, emulate_bf16_storage
hp = dict(find_model_spec(load_parameters('conf/base/parameters.yml'), 'image_classifier'))
hp['architecture'] = copy.deepcopy(hp['architecture']); hp['architecture'][-1]['fully_connected']['out_features'] = 10
def rel(a, b): return float((a.float() - b.float()).abs().max() / max(float(b.abs().max()), 1e-12))
for batch in (8, 128):
    torch.manual_seed(1)
    oracle = OracleDeepcvModule((3, 32, 32), hp)
    em = emulate_bf16_storage(copy.deepcopy(oracle))
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, 3, 32, 32, generator=g).bfloat16().float(); y = torch.randint(0, 10, (batch,), generator=g)
    l1, lo1 = train_step(oracle, x, y); l2, lo2 = train_step(em, x, y)
    print(batch, 'logits', rel(lo2, lo1), 'loss', abs(l1-l2)/l1)
    print({n.replace('_child_modules.', '').replace('_submodule_', 's'): f'{rel(p.grad, dict(oracle.named_parameters())[n].grad):.1e}' for n, p in em.named_parameters()})
