import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deepcv_b200.meta.data.datasets import dataloader_prefetch_batches
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(3)
batches = [(torch.randint(0, 256, (64, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory(), torch.randint(0, 10, (64,), generator=g).pin_memory()) for _ in range(7)]
batches.append((torch.randint(0, 256, (5, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory(), torch.randint(0, 10, (5,), generator=g).pin_memory()))
loader = dataloader_prefetch_batches(batches, dev)
busy = torch.randn(2048, 2048, device=dev)
labels, xs = [], []
for x, y in loader:
    for _ in range(3):
        busy = (busy @ busy).clamp_(-1, 1)
    xs.append(x.clone()); labels.append(y.clone())
torch.cuda.synchronize()
for i, (l, xx, b) in enumerate(zip(labels, xs, batches)):
    print(i, torch.equal(l.cpu(), b[1]), torch.equal(xx.cpu(), b[0]), l.shape, l.dtype, b[1].dtype, l[:5].tolist(), b[1][:5].tolist())
