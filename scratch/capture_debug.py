""" Bisect the capture failure: each variant in its own process. """
import os, subprocess, sys
VARIANTS = ['base', 'noarena', 'nodone', 'nozero', 'fp32cast']
if len(sys.argv) == 1:
    for v in VARIANTS:
        r = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True)
        tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
        print(f'=== {v}: rc={r.returncode}'); print('\n'.join(tail))
    sys.exit(0)
variant = sys.argv[1]
sys.path.insert(0, '/root/repo')
import torch
from deepcv_b200 import ops
from deepcv_b200.meta.base_module import DeepcvModule
from deepcv_b200.meta import flat_params as FP
from deepcv_b200.meta.flat_params import FlatAdamW, flatten_parameters
from deepcv_b200.meta.ignite_training import CrossEntropyLoss, GraphedTrainStep
from deepcv_b200.yaml_config import benchmark_model_spec
hp = benchmark_model_spec('/root/repo/conf/base/parameters.yml', 'image_classifier', out_features=10)
dev = torch.device('cuda', 0)
torch.manual_seed(0)
m = DeepcvModule((3, 32, 32), hp).to(dev)
opt = FlatAdamW(m.parameters(), lr=1e-3).attach(flatten_parameters(m))
x = torch.randn(8, 3, 32, 32, device=dev); y = torch.randint(0, 10, (8,), device=dev)
if variant == 'nodone':
    ops._backward_done = lambda g, t: None
if variant == 'nozero':
    FP.FlatParameters.reset_gradients = lambda self, memset=True: self.restore_grad_views()
orig_bwd = torch.Tensor.backward
def traced_backward(self, *a, **k):
    print('backward: capturing =', torch.cuda.is_current_stream_capturing(), 'stream', torch.cuda.current_stream().cuda_stream, flush=True)
    return orig_bwd(self, *a, **k)
torch.Tensor.backward = traced_backward
# trace every library call made while capturing, with thread and stream
import threading
from deepcv_b200 import _lib
real = _lib.lib._load()
class Tracer:
    def __getattr__(self, name):
        fn = getattr(real, name)
        def w(*a):
            if torch.cuda.is_current_stream_capturing():
                print(f'  [{threading.current_thread().name}] {name} stream={torch.cuda.current_stream().cuda_stream:#x}', flush=True)
            return fn(*a)
        return w
tr = Tracer()
ops.lib = tr; FP.lib = tr
import deepcv_b200.meta.ignite_training as IT
IT.lib = tr
r = GraphedTrainStep(m, CrossEntropyLoss(), opt, x, y, warmup_iters=2, use_accumulator_arena=(variant != 'noarena'))
r.step(x, y); torch.cuda.synchronize()
print('OK', float(r.static_loss))
