import copy, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import _make_block
dev = torch.device('cuda')
def rel(a, b): return float((a.detach().double().cpu() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))
for (n, c, h, k, act, bn) in [(6,128,12,128,'leaky',True), (6,128,12,128,'leaky',False), (6,128,12,128,'none',False), (6,64,24,64,'leaky',True), (6,256,6,256,'leaky',True), (2,128,12,128,'leaky',True), (6,128,12,64,'leaky',True), (6,64,12,128,'leaky',True), (6,128,16,128,'leaky', True), (6,128,8,128,'leaky', True)]:
    ref, ours = _make_block(c, k, (3,3), (1,1), (1,1), (1,1), act, bn, 0, seed=1)
    ours = ours.to(dev)
    x = torch.randn(n, c, h, h)
    r64 = copy.deepcopy(ref).double()
    xr = x.double().requires_grad_(True); yr = r64(xr); dy = torch.randn(yr.shape); yr.backward(dy.double())
    xd = x.to(dev).requires_grad_(True); y = ours(xd); y.backward(dy.to(dev))
    print((n,c,h,k,act,bn), 'y', f'{rel(y, yr):.1e}', 'dx', f'{rel(xd.grad, xr.grad):.1e}', {nm: f'{rel(p.grad, q.grad):.1e}' for (nm, p), (_, q) in zip(ours.named_parameters(), r64.named_parameters())})
