import copy, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import _small_resnet_hp, _run_model
from deepcv_b200 import ops
from deepcv_b200.meta.base_module import DeepcvModule
from deepcv_b200.meta.ignite_training import CrossEntropyLoss
from oracle.deepcv_oracle import OracleDeepcvModule, train_step
dev = torch.device('cuda')
hp = _small_resnet_hp()
torch.manual_seed(11)
init = OracleDeepcvModule((3, 96, 96), hp)
model = DeepcvModule((3, 96, 96), hp); model.load_state_dict(init.state_dict()); model = model.to(dev)
g = torch.Generator().manual_seed(5)
x = torch.randn(6, 3, 96, 96, generator=g); y = torch.randint(0, 17, (6,), generator=g)
o64 = copy.deepcopy(init).double()
cap = {}
bb = o64._child_modules['_submodule_0']._child_modules
for name in ('_submodule_13', '_submodule_14', '_submodule_22'):
    blk = bb[name]
    def mk(name):
        def fh(mod, inp, out):
            cap[name + '.pre'] = out
            out.register_hook(lambda gr: cap.__setitem__(name + '.dy', gr))
        return fh
    blk[0].register_forward_hook(mk(name))
    def mk2(name):
        def fh(mod, inp, out):
            cap[name + '.z'] = out
            out.register_hook(lambda gr: cap.__setitem__(name + '.dz', gr))
        return fh
    blk.register_forward_hook(mk2(name))
    blk[1].register_forward_hook(lambda mod, inp, out, name=name: cap.__setitem__(name + '.y', out))
train_step(o64, x.double(), y)
ops._DEBUG_CAPTURE = []
_run_model(model, x.to(dev), y.to(dev), CrossEntropyLoss())
caps = ops._DEBUG_CAPTURE  # in backward order: last layer first
def rel(a, b): return float((a.detach().double().cpu() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))
names = [n for n, m in model._submodules['_submodule_0']._submodules.items() if hasattr(m, '_op')]
order = list(reversed(names))
for nm, c in zip(order, caps):
    if nm in ('_submodule_13', '_submodule_14', '_submodule_22'):
        print(nm, c['wshape'], 'dz', rel(c['dz'], cap[nm + '.dz']), 'y', rel(c['y'], cap[nm + '.y']), 'dy', rel(c['dy'], cap[nm + '.dy']))
        # per-channel error of dy
        e = (c['dy'].double().cpu() - cap[nm + '.dy']).abs().amax(dim=(0, 2, 3)); sc = cap[nm + '.dy'].abs().amax(dim=(0, 2, 3))
        worst = torch.argsort(e / sc, descending=True)[:5]
        print('   worst channels', worst.tolist(), (e / sc)[worst].tolist())
        yy = cap[nm + '.y']; mu = yy.mean(dim=(0, 2, 3)); sd = yy.std(dim=(0, 2, 3))
        print('   mu/sd of those', (mu / sd)[worst].tolist(), 'sd', sd[worst].tolist())
o64p = dict(o64.named_parameters())
for n, p in model.named_parameters():
    if any(t in n for t in ('_submodule_13.', '_submodule_14.', '_submodule_22.')):
        print(n[-30:], f'{rel(p.grad, o64p[n].grad):.2e}')
for nm, c in zip(order, caps):
    if nm == '_submodule_14':
        ref_db = cap[nm + '.dy'].sum(dim=(0, 2, 3))
        print('dbias from captured dy', rel(c['dy'].double().sum(dim=(0, 2, 3)), ref_db))
        print('dw captured vs ref', rel(c['dw'], o64p['_child_modules._submodule_0._child_modules._submodule_14.0.weight'].grad))
        xx = c['x'].double().cpu(); dyy = c['dy'].double().cpu()
        dw_from_cap = torch.nn.grad.conv2d_weight(xx.contiguous(), (128, 128, 3, 3), dyy.contiguous(), padding=1)
        print('dw recomputed from captured x,dy vs ref', rel(dw_from_cap, o64p['_child_modules._submodule_0._child_modules._submodule_14.0.weight'].grad))
for nm, c in zip(order, caps):
    if nm in ('_submodule_14', '_submodule_22'):
        dz, yy, pqr, dy = c['dz'].double().cpu(), c['y'].double().cpu(), c['pqr'].double().cpu(), c['dy'].double().cpu()
        P, Q, R = (pqr[:, :, i][:, :, None, None] for i in range(3))
        pre = P * dz + Q * yy + R
        exp = torch.where(yy > 0, pre, 0.01 * pre)
        print(nm, 'dy vs formula(captured dz,y,pqr)', rel(dy, exp), ' formula vs ref', rel(exp, cap[nm + '.dy']))
        print('   ptrs', hex(c['dz_ptr']), hex(c['y_ptr']), hex(c['dy_ptr']), 'sizes', dz.numel() * 4)
        # expected P,Q,R from the oracle's BN
        bn = bb[nm][2]
        print('   pqr[0,:3]', pqr[0, :3].tolist())
for nm, c in zip(order, caps):
    if nm in ('_submodule_14', '_submodule_22'):
        bn = copy.deepcopy(bb[nm][2]); bn.train()
        y_ = cap[nm + '.y'].detach().clone().requires_grad_(True)
        z_ = bn(y_); z_.backward(cap[nm + '.dz'])
        pre = cap[nm + '.pre'].detach()
        dy_f = y_.grad * torch.where(pre > 0, torch.ones_like(pre), 0.01 * torch.ones_like(pre))
        print(nm, 'oracle dy vs torch-BN-bwd(oracle dz, y):', rel(cap[nm + '.dy'], dy_f), ' ours vs that:', rel(c['dy'], dy_f))
        print('   z consistency', rel(z_, cap[nm + '.z']), 'training flag', bb[nm][2].training)
        print('   again: ours vs cap.dy', rel(c['dy'], cap[nm + '.dy']), ' cap.dy vs dy_f', rel(cap[nm + '.dy'], dy_f), 'shapes', c['dy'].shape, cap[nm + '.dy'].shape, dy_f.shape, 'max', float(cap[nm + '.dy'].abs().max()), float(dy_f.abs().max()), float(c['dy'].abs().max()))
        d = (c['dy'].double().cpu() - cap[nm + '.dy']).abs(); idx = torch.nonzero(d == d.max())[0].tolist(); print('   argmax', idx, float(c['dy'].double().cpu()[tuple(idx)]), float(cap[nm + '.dy'][tuple(idx)]), float(dy_f[tuple(idx)]), 'pre', float(pre[tuple(idx)]), 'y ours', float(c['y'].cpu()[tuple(idx)]))
