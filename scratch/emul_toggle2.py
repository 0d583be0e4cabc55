import copy, sys, torch
sys.path.insert(0, '.')
import oracle.deepcv_oracle as O
import torch.nn.functional as F
from deepcv_b200.yaml_config import find_model_spec, load_parameters
hp = dict(find_model_spec(load_parameters('conf/base/parameters.yml'), 'image_classifier'))
hp['architecture'] = copy.deepcopy(hp['architecture']); hp['architecture'][-1]['fully_connected']['out_features'] = 10
def rel(a, b): return float((a.float() - b.float()).abs().max() / max(float(b.abs().max()), 1e-12))
torch.manual_seed(1)
oracle = O.OracleDeepcvModule((3, 32, 32), hp)
g = torch.Generator().manual_seed(2)
x = torch.randn(128, 3, 32, 32, generator=g).bfloat16().float(); y = torch.randint(0, 10, (128,), generator=g)
O.train_step(oracle, x, y)
ref = {n: p.grad.clone() for n, p in oracle.named_parameters()}
q = lambda t: t.to(torch.bfloat16).float() + (t - t.detach()) * 0  # placeholder
def Q(t): return O._Round.apply(t, True, False)
for mode in ('w', 'y', 'z', 'pool'):
    em = copy.deepcopy(oracle)
    for m in em.modules():
        if isinstance(m, torch.nn.Sequential) and any(isinstance(c, torch.nn.Conv2d) for c in m):
            def fwd(x, m=m, mode=mode):
                op = m[0]
                w = Q(op.weight) if mode == 'w' else op.weight
                t = m[1](F.conv2d(x, w, op.bias, op.stride, op.padding))
                if mode == 'y': t = Q(t)
                for n in list(m)[2:]: t = n(t)
                if mode == 'z': t = Q(t)
                return t
            m.forward = fwd
        elif isinstance(m, torch.nn.AvgPool2d) and mode == 'pool':
            orig = m.forward
            m.forward = lambda x, orig=orig: Q(orig(x))
    O.train_step(em, x, y)
    errs = {n: rel(p.grad, ref[n]) for n, p in em.named_parameters() if not n.endswith('2.bias')}
    v = sorted(errs.values())
    print(mode, 'max', v[-1], 'median', v[len(v)//2])
