#!/usr/bin/env python
""" Per-layer convolution timing on the B200 through the C ABI (CUDA events, working set cycled beyond L2): forward / data gradient /
weight gradient of the ResNet-style (C4) layer shapes, in TFLOP/s and as a fraction of the measured dense bf16 peak (MEASURED_PEAKS.json). """
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deepcv_b200._lib import ACT_LEAKY_RELU, ALGO_AUTO, ALGO_DIRECT, DCV_BF16, ConvShape, check, lib  # noqa: E402

LAYERS = [  # name, c, h, k, ksize
    ('s1 64->64 @56', 64, 56, 64, 3), ('s2in 64->128 @28', 64, 28, 128, 3), ('s2 128->128 @28', 128, 28, 128, 3), ('s3in 128->256 @14', 128, 14, 256, 3),
    ('s3 256->256 @14', 256, 14, 256, 3), ('s4in 256->512 @7', 256, 7, 512, 3), ('s4 512->512 @7', 512, 7, 512, 3),
]


STEM_LAYERS = [('stem 3->64 7x7/2 @224', 3, 224, 64, 7, 2)]   # the gather kernels (no data gradient: the input is the image)
CIFAR_LAYERS = [('c1 3->4 5x5 @32', 3, 32, 4, 5), ('c2 4->4 5x5 @32', 4, 32, 4, 5), ('c4 4->16 3x3 @16', 4, 16, 16, 3), ('c5 16->16 3x3 @16', 16, 16, 16, 3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--net', default='resnet', choices=['resnet', 'cifar', 'stem'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--ops', default='fwd,dgrad,wgrad')
    ap.add_argument('--algo', default='auto')
    ap.add_argument('--only', default=None, help='substring filter on the layer name')
    args = ap.parse_args()
    algo = ALGO_DIRECT if args.algo == 'direct' else ALGO_AUTO
    peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()) if (ROOT / 'MEASURED_PEAKS.json').exists() else {'bf16_tflops': 1590.0}
    dev = torch.device('cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    n = args.batch
    for name, c, h, k, ks, *rest in {'resnet': LAYERS, 'cifar': CIFAR_LAYERS, 'stem': STEM_LAYERS}[args.net]:
        if args.only and args.only not in name:
            continue
        pad, stride = ks // 2, (rest[0] if rest else 1)
        ho = (h + 2 * pad - ks) // stride + 1
        shape = ConvShape(n, h, h, c, k, ks, ks, stride, stride, pad, pad, 1, 1, ho, ho)
        act_bytes = n * (h * h * c + ho * ho * k) * 2
        reps = max(2, int(300e6 // act_bytes) + 1)
        xs = [torch.randn(n, h, h, c, device=dev).bfloat16() for _ in range(reps)]
        ys = [torch.randn(n, ho, ho, k, device=dev).bfloat16() for _ in range(reps)]
        w = (torch.randn(k, ks, ks, c, device=dev) * 0.05).bfloat16()
        w32 = w.float()
        wt = torch.empty(c, ks, ks, k, device=dev, dtype=torch.bfloat16)
        check(lib.dcv_pack_conv_weight(P(w32), P(wt), DCV_BF16, k, ks, ks, c, 1, st), 'pack')
        bias = torch.zeros(k, device=dev)
        dw = torch.empty(k, ks, ks, c, device=dev)
        ws_bytes = int(lib.dcv_conv2d_wgrad_workspace(ctypes.byref(shape), DCV_BF16, algo))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        flop = 2.0 * n * ho * ho * k * c * ks * ks
        row = {'layer': name}
        gather = None
        if stride != 1:   # served by the gather kernels (software im2col in shared memory): their own weight layout
            sc = ks * c
            kpad = (ks * ((sc + 7) // 8 * 8) + 63) // 64 * 64
            assert lib.dcv_conv2d_gather_supported(ctypes.byref(shape), P(xs[0]), kpad, DCV_BF16)
            w_col = torch.empty(k, kpad, device=dev, dtype=torch.bfloat16)
            check(lib.dcv_gather_pack_weight(P(w), P(w_col), k, ks, sc, kpad, DCV_BF16, st), 'gather_pack_weight')
            gather = (w_col, kpad, torch.empty(k, kpad, device=dev))
            if args.algo != 'gather' and lib.dcv_conv2d_pairs_supported(ctypes.byref(shape), P(xs[0]), DCV_BF16):   # stride 2, <= 4 channels: the pixel-pair kernels
                w_col = torch.empty(k, 256, device=dev, dtype=torch.bfloat16)
                check(lib.dcv_pairs_pack_weight(P(w), P(w_col), ctypes.byref(shape), DCV_BF16, st), 'pairs_pack_weight')
                gather = (w_col, 0, torch.empty(k, 256, device=dev))
        for op in args.ops.split(','):
            if op == 'dgrad' and stride != 1:
                continue
            def launch(i):
                if gather is not None and gather[1] == 0 and op == 'fwd':
                    check(lib.dcv_conv2d_fwd_pairs(ctypes.byref(shape), P(xs[i]), P(gather[0]), P(bias), P(ys[i]), None, ACT_LEAKY_RELU, 0.01, 0, st), op)
                elif gather is not None and gather[1] == 0:
                    check(lib.dcv_conv2d_wgrad_pairs(ctypes.byref(shape), P(xs[i]), P(ys[i]), P(gather[2]), 0, st), op)
                elif gather is not None and op == 'fwd':
                    check(lib.dcv_conv2d_fwd_gather(ctypes.byref(shape), P(xs[i]), P(gather[0]), gather[1], P(bias), P(ys[i]), None, ACT_LEAKY_RELU, 0.01, 0, st), op)
                elif gather is not None:
                    check(lib.dcv_conv2d_wgrad_gather(ctypes.byref(shape), P(xs[i]), P(ys[i]), P(gather[2]), gather[1], 0, st), op)
                elif op == 'fwd':
                    check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xs[i]), P(w), P(bias), P(ys[i]), None, ACT_LEAKY_RELU, 0.01, DCV_BF16, algo, 0, st), op)
                elif op == 'dgrad':
                    check(lib.dcv_conv2d_dgrad(ctypes.byref(shape), P(ys[i]), P(w), P(wt), P(xs[i]), DCV_BF16, algo, st), op)
                else:
                    check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xs[i]), P(ys[i]), P(dw), P(ws), DCV_BF16, algo, 0, st), op)
            launch(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                for i in range(reps):
                    launch(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (args.iters * reps)
            row[op] = dict(ms=round(ms, 4), tflops=round(flop / ms / 1e9, 1), frac_of_peak=round(flop / ms / 1e9 / peaks['bf16_tflops'], 3),
                           tc=int(gather is not None or lib.dcv_conv2d_tc_supported(ctypes.byref(shape), DCV_BF16, {'fwd': 0, 'dgrad': 1, 'wgrad': 2}[op])) if algo == ALGO_AUTO else 0)
        print(json.dumps(row))


if __name__ == '__main__':
    main()
