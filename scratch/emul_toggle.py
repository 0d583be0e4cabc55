import copy, sys, torch
sys.path.insert(0, '.')
import oracle.deepcv_oracle as O
from deepcv_b200.yaml_config import find_model_spec, load_parameters
hp = dict(find_model_spec(load_parameters('conf/base/parameters.yml'), 'image_classifier'))
hp['architecture'] = copy.deepcopy(hp['architecture']); hp['architecture'][-1]['fully_connected']['out_features'] = 10
def rel(a, b): return float((a.float() - b.float()).abs().max() / max(float(b.abs().max()), 1e-12))
orig_q = O._q
torch.manual_seed(1)
oracle = O.OracleDeepcvModule((3, 32, 32), hp)
g = torch.Generator().manual_seed(2)
x = torch.randn(128, 3, 32, 32, generator=g).bfloat16().float(); y = torch.randint(0, 10, (128,), generator=g)
O.train_step(oracle, x, y)
ref = {n: p.grad.clone() for n, p in oracle.named_parameters()}
for mode in ('all', 'fwd_only', 'bwd_only', 'none'):
    def q(x, fwd=True, bwd=False, mode=mode):
        if mode == 'fwd_only': bwd = False
        if mode == 'bwd_only': fwd = False
        if mode == 'none': fwd = bwd = False
        return orig_q(x, fwd, bwd)
    O._q = q
    em = O.emulate_bf16_storage(copy.deepcopy(oracle))
    O.train_step(em, x, y)
    errs = [rel(p.grad, ref[n]) for n, p in em.named_parameters() if not n.endswith('2.bias')]
    print(mode, 'max', max(errs), 'median', sorted(errs)[len(errs)//2])
