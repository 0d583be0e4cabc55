#!/usr/bin/env python
""" HBM-bound kernels of the path through the C ABI (CUDA events, buffers cycled beyond L2): achieved GB/s on ALGORITHMIC bytes vs the measured
copy bandwidth (MEASURED_PEAKS.json). Covers the BatchNorm/GroupNorm streaming kernels, pooling, links, im2col and the uint8 preprocess sweep
(BASELINE.json configs[4]: 32^2 .. 1024^2). """
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deepcv_b200._lib import ACT_LEAKY_RELU, DCV_BF16, DCV_F32, ConvShape, check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


ITERS = 10


WORK_STREAM = None


def timeit(fn, reps, iters=None):
    """ ms per launch. The `reps` launches (distinct buffers, working set > L2) are captured into ONE CUDA graph and the graph is replayed: the
    kernels are 5-50 us, shorter than a Python + ctypes launch (memset + kernel), so eager launches would time the host, not the device. """
    iters = ITERS if iters is None else iters
    fn(0)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=WORK_STREAM):
        for i in range(reps):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(WORK_STREAM):
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='norm,pool,preprocess,im2col')
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--sizes', default='32,64,128,224,256,512,1024')
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    global ITERS
    ITERS = args.iters
    peak = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())['hbm_gbs'] if (ROOT / 'MEASURED_PEAKS.json').exists() else 6650.0
    dev = torch.device('cuda')
    global WORK_STREAM
    WORK_STREAM = torch.cuda.Stream()
    torch.cuda.set_stream(WORK_STREAM)
    st = ctypes.c_void_p(WORK_STREAM.cuda_stream)
    what = args.what.split(',')

    def report(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        print(json.dumps(dict(kernel=name, us=round(ms * 1e3, 2), algorithmic_MB=round(nbytes / 1e6, 2), GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / peak, 3))))

    if 'norm' in what or 'pool' in what:
        for (c, hw, dt, tdt, es) in [(64, 56 * 56, DCV_BF16, torch.bfloat16, 2), (128, 28 * 28, DCV_BF16, torch.bfloat16, 2), (512, 49, DCV_BF16, torch.bfloat16, 2), (4, 1024, DCV_BF16, torch.bfloat16, 2), (16, 256, DCV_BF16, torch.bfloat16, 2),
                                 (64, 56 * 56, DCV_F32, torch.float32, 4)]:
            n = args.batch if c >= 64 else 512
            E = n * hw * c
            reps = max(2, int(400e6 // (E * es * 2)) + 1)
            ys = [torch.randn(n, hw, c, device=dev).to(tdt) for _ in range(reps)]
            zs = [torch.empty_like(y) for y in ys]
            dz = [torch.randn(n, hw, c, device=dev).to(tdt) for _ in range(min(reps, 4))]
            stats = torch.empty(n, c, 2, device=dev)
            ab = torch.rand(n, c, 2, device=dev)
            pqr = torch.rand(n, c, 3, device=dev)
            dbias = torch.empty(c, device=dev)
            tag = f'c{c} hw{hw} n{n} {"bf16" if es == 2 else "fp32"}'
            if 'norm' in what:
                report(f'stats_kernel {tag}', timeit(lambda i: check(lib.dcv_norm_stats(P(ys[i]), P(stats), n, hw, c, dt, 0, st)), reps), E * es)
                report(f'apply_fwd_kernel {tag}', timeit(lambda i: check(lib.dcv_norm_apply_fwd(P(ys[i]), P(ab), P(zs[i]), n, hw, c, dt, st)), reps), 2 * E * es)
                report(f'bwd_reduce_kernel {tag}', timeit(lambda i: check(lib.dcv_norm_bwd_reduce(P(dz[i % len(dz)]), P(ys[i]), P(stats), n, hw, c, dt, 0, st)), reps), 2 * E * es)
                report(f'bwd_apply_kernel {tag}', timeit(lambda i: check(lib.dcv_act_norm_bwd_apply(P(dz[i % len(dz)]), P(ys[i]), P(pqr), P(zs[i]), P(dbias), ACT_LEAKY_RELU, 0.01, n, hw, c, dt, 0, st)), reps), 3 * E * es)
            if 'pool' in what and hw in (3136, 784, 1024, 256):
                h = int(hw ** 0.5)
                outs = [torch.empty(n, h // 2, h // 2, c, device=dev, dtype=tdt) for _ in range(reps)]
                report(f'avgpool_fwd 2x2 {tag}', timeit(lambda i: check(lib.dcv_avgpool2d_fwd(P(ys[i]), P(outs[i]), n, h, h, c, 2, 2, 2, 2, dt, st)), reps), int(1.25 * E * es))
                report(f'avgpool_bwd 2x2 {tag}', timeit(lambda i: check(lib.dcv_avgpool2d_bwd(P(outs[i]), P(zs[i]), n, h, h, c, 2, 2, 2, 2, dt, st)), reps), int(1.25 * E * es))
                report(f'axpby (residual sum) {tag}', timeit(lambda i: check(lib.dcv_axpby(P(ys[i]), P(dz[i % len(dz)]), P(zs[i]), 1., 1., E, dt, st)), reps), 3 * E * es)
            del ys, zs, dz
    if 'preprocess' in what:
        mean = torch.tensor([0.485, 0.456, 0.406], device=dev)
        std = torch.tensor([0.229, 0.224, 0.225], device=dev)
        for size in [int(v) for v in args.sizes.split(',')]:
            for out_dt, tdt, es in ((DCV_BF16, torch.bfloat16, 2), (DCV_F32, torch.float32, 4)):
                n = max(4, int(1024e6 // (size * size * 3 * (1 + es))))
                n = min(n, 262144)
                reps = 3
                pad = size // 8
                imgs = [torch.randint(0, 256, (n, size, size, 3), device=dev, dtype=torch.uint8) for _ in range(reps)]
                outs = [torch.empty(n, size, size, 3, device=dev, dtype=tdt) for _ in range(reps)]
                flip = (torch.rand(n, device=dev) < 0.5).to(torch.uint8)
                crop = torch.randint(0, 2 * pad + 1, (n, 2), device=dev, dtype=torch.int32)
                ms = timeit(lambda i: check(lib.dcv_preprocess_u8(P(imgs[i]), P(outs[i]), n, size, size, 3, size, size, pad, P(mean), P(std), P(flip), P(crop), out_dt, 3, 0, st)), reps)
                report(f'preprocess_u8 {size}x{size} n{n} -> {"bf16" if es == 2 else "fp32"} (random flip + pad-crop)', ms, n * size * size * 3 * (1 + es))
                del imgs, outs
    if 'im2col' in what:
        n = args.batch
        shape = ConvShape(n, 224, 224, 3, 64, 7, 7, 2, 2, 3, 3, 1, 1, 112, 112)
        x = torch.randn(n, 224, 224, 3, device=dev).bfloat16()
        col = torch.empty(n, 112, 112, 192, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda i: check(lib.dcv_im2col(ctypes.byref(shape), P(x), P(col), 192, DCV_BF16, st)), 1)
        report(f'im2col stem 3->kpad192 n{n}', ms, col.numel() * 2 + x.numel() * 2)


if __name__ == '__main__':
    main()
