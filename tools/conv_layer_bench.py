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


CIFAR_LAYERS = [('c1 3->4 5x5 @32', 3, 32, 4, 5), ('c2 4->4 5x5 @32', 4, 32, 4, 5), ('c4 4->16 3x3 @16', 4, 16, 16, 3), ('c5 16->16 3x3 @16', 16, 16, 16, 3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--net', default='resnet', choices=['resnet', 'cifar'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--ops', default='fwd,dgrad,wgrad')
    ap.add_argument('--algo', default='auto')
    ap.add_argument('--only', default=None, help='substring filter on the layer name')
    args = ap.parse_args()
    algo = ALGO_AUTO if args.algo == 'auto' else ALGO_DIRECT
    peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()) if (ROOT / 'MEASURED_PEAKS.json').exists() else {'bf16_tflops': 1590.0}
    dev = torch.device('cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    n = args.batch
    for name, c, h, k, ks in (LAYERS if args.net == 'resnet' else CIFAR_LAYERS):
        if args.only and args.only not in name:
            continue
        pad = ks // 2
        shape = ConvShape(n, h, h, c, k, ks, ks, 1, 1, pad, pad, 1, 1, h, h)
        act_bytes = n * h * h * (c + k) * 2
        reps = max(2, int(300e6 // act_bytes) + 1)
        xs = [torch.randn(n, h, h, c, device=dev).bfloat16() for _ in range(reps)]
        ys = [torch.randn(n, h, h, k, device=dev).bfloat16() for _ in range(reps)]
        w = (torch.randn(k, ks, ks, c, device=dev) * 0.05).bfloat16()
        w32 = w.float()
        wt = torch.empty(c, ks, ks, k, device=dev, dtype=torch.bfloat16)
        check(lib.dcv_pack_conv_weight(P(w32), P(wt), DCV_BF16, k, ks, ks, c, 1, st), 'pack')
        bias = torch.zeros(k, device=dev)
        dw = torch.empty(k, ks, ks, c, device=dev)
        ws_bytes = int(lib.dcv_conv2d_wgrad_workspace(ctypes.byref(shape), DCV_BF16, algo))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        flop = 2.0 * n * h * h * k * c * ks * ks
        row = {'layer': name}
        for op in args.ops.split(','):
            def launch(i):
                if op == 'fwd':
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
                           tc=int(lib.dcv_conv2d_tc_supported(ctypes.byref(shape), DCV_BF16, {'fwd': 0, 'dgrad': 1, 'wgrad': 2}[op])) if algo == ALGO_AUTO else 0)
        print(json.dumps(row))


if __name__ == '__main__':
    main()
