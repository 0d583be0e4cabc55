#!/usr/bin/env python
""" Benchmark of the DeepcvModule conv/BN/augment hot path (BASELINE.json): train images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cifar|imagenet|preprocess] [--impl reference]

A step = fused uint8 preprocess (normalise / flip / crop) -> forward -> cross-entropy -> backward -> (bucketed NCCL gradient all-reduce when
N > 1) -> AdamW, on one synthetic batch per GPU (weak scaling). BASELINE.json's metric names two shapes and ONE invocation measures both
(`--workload all`, the default): the headline keys of the line are configs[1] — the default CIFAR-10 `image_classifier` DeepcvModule, bf16, batch
512 per GPU — and `workloads.imagenet` carries configs[3] — the ResNet-style DeepcvModule at 3x224x224, bf16, batch 256 per GPU — with its own
value / e2e / ms_per_step / clocks / tensor roofline (per-layer forward / data-gradient / weight-gradient fractions of the measured bf16 peak).
Both are stepped through the public API: `ignite_training.make_process_function` (what `train()` wires into the `Engine`), which captures the step
into a CUDA graph on its first call. Prints ONE JSON line (rank 0):
  value       images/s with the uint8 batches already resident in HBM (a pool of distinct batches larger than L2 is cycled)
  e2e         same metric through the public API with HOST (pinned) uint8 batches: H2D copy of images / labels / augmentation parameters and
              D2H read of the loss inside the timed region, every step
  roofline    the dominant kernel of the step, timed with CUDA events in this process
  cpu_baseline  the reference CPU path (the oracle: stock torch CPU modules, fp32) timed on this box's host cores on a bounded sample
`--impl reference` times only that CPU path and prints the same line with "impl": "reference".
"""
import argparse
import copy
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC, UNIT = 'train images/sec', 'images/s'
CIFAR_MEAN, CIFAR_STD = [0.491, 0.482, 0.447], [0.247, 0.243, 0.261]
IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None, help='timed steps of each workload (default: 1000 cifar, 50 imagenet: about half a second of device time)')
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='all', choices=['all', 'cifar', 'imagenet', 'preprocess'])
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (default: 512 cifar, 256 imagenet)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='budget of the bounded CPU baseline sample')
    ap.add_argument('--no-layer-table', action='store_true', help='skip the per-layer convolution table of the imagenet workload')
    return ap.parse_args()


def steps_for(args, workload: str) -> int:
    return args.steps if args.steps is not None else (50 if workload == 'imagenet' else 1000)


def workload_spec(name: str, batch):
    from deepcv_b200.yaml_config import benchmark_model_spec
    if name == 'cifar':
        hp = benchmark_model_spec(ROOT / 'conf' / 'base' / 'parameters.yml', 'image_classifier')
        classes, size, mean, std, pad, b = 10, 32, CIFAR_MEAN, CIFAR_STD, 4, 512
        label = 'deepcv.classification.image default image_classifier DeepcvModule (conf/base/parameters.yml), synthetic CIFAR-10-shaped uint8 3x32x32'
    else:
        hp = benchmark_model_spec(ROOT / 'conf' / 'base' / 'resnet_style.yml', 'resnet_style_classifier')
        classes, size, mean, std, pad, b = 1000, 224, IMAGENET_MEAN, IMAGENET_STD, 16, 256
        label = 'ResNet-style DeepcvModule with residual/dense links (conf/base/resnet_style.yml), synthetic ImageNet-shaped uint8 3x224x224'
    hp['architecture'][-1]['fully_connected']['out_features'] = classes
    return dict(hp=hp, classes=classes, size=size, mean=mean, std=std, pad=pad, batch=batch or b, label=label)


def needs_remeasure(clocks: dict) -> bool:
    """ Timing rule: a run that saw hw_slowdown / hw_thermal_slowdown / sw_thermal_slowdown, or SM clocks stuck well below max with no reason at all, is
    rejected and re-measured once; sw_power_cap is kept and noted. """
    bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    if bad & set(clocks.get('reasons') or []):
        return True
    sm, mx = clocks.get('sm_mhz'), clocks.get('sm_max_mhz')
    return bool(sm and mx and not clocks.get('reasons') and sm < 0.7 * mx)


class ClockSampler:
    """ nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). The sampler is started early and
    `wait_ready()` blocks until it has delivered its first row, so that even a 100 ms timed region is covered; rows carry the host time at
    which they were read and `summary(t0, t1)` keeps the ones inside the region (or the nearest one when the region is shorter than a period). """
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int, period_ms: int = 20):
        self.rows, self.proc, self.index, self.period_ms = [], None, index, period_ms

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def wait_ready(self, timeout: float = 10.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(2.5 * self.period_ms / 1e3)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self, t0=None, t1=None):
        rows = [(t, r) for t, r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        if t0 is not None and rows:
            inside = [(t, r) for t, r in rows if t0 <= t <= t1 + 1.5 * self.period_ms / 1e3]
            rows = inside or [min(rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))]
        sm = [float(r[0]) for _, r in rows]
        mx = [float(r[1]) for _, r in rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for _, r in rows for i in range(4) if r[3 + i].lower().startswith('active')})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def cpu_reference_run(spec, steps: int, warmup: int, seconds: float, batch: int):
    """ The reference CPU path: oracle DeepcvModule (stock torch CPU modules, fp32) + torch.optim.AdamW + the ToTensor/Normalize/crop/flip
    restatement, on host cores. Sweeps thread counts, keeps the best; each step is a bounded sample (`batch` images). """
    import torch
    from oracle import deepcv_oracle as O
    torch.manual_seed(563454)
    model = O.OracleDeepcvModule((3, spec['size'], spec['size']), spec['hp'])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    g = torch.Generator().manual_seed(434546)
    img = torch.randint(0, 256, (batch, spec['size'], spec['size'], 3), generator=g, dtype=torch.uint8)
    y = torch.randint(0, spec['classes'], (batch,), generator=g)
    cores = os.cpu_count() or 1
    best = None
    candidates = sorted({max(1, cores // 2), cores}) if cores > 2 else [cores]
    for k in candidates:
        torch.set_num_threads(k)
        times = []
        t_start = time.perf_counter()
        for i in range(warmup + steps):
            flip, crop = O.draw_augmentation_params(batch, spec['pad'], 434546 + i)
            t0 = time.perf_counter()
            x = O.preprocess_u8(img, spec['mean'], spec['std'], flip=flip, crop_yx=crop, pad=spec['pad'])
            O.train_step(model, x, y, optimizer=opt)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_start > seconds / len(candidates) and len(times) >= 3:
                break
        med = statistics.median(times)
        if best is None or med < best['ms'] / 1e3:
            best = dict(ms=med * 1e3, threads=k, steps=len(times))
    del model, opt
    return dict(value=batch / (best['ms'] / 1e3), unit=UNIT, cores=best['threads'], kind='port', ms_per_step=best['ms'],
                sample=f"{best['steps']} timed steps of batch {batch} (fp32, preprocess + forward + backward + AdamW), best of thread counts {candidates} on {cores} host cores; median step")


def run_reference(args):
    """ The reference arm: the reference's CPU implementation of the path (oracle port: the reference package cannot be imported here, SURVEY.md section
    8.c) on this box's host cores, same metric / config, each step a bounded sample. Rank 0 alone works. """
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    names = ['cifar', 'imagenet'] if args.workload in ('all', 'preprocess') else [args.workload]
    results = {}
    for name in names:
        spec = workload_spec(name, args.batch)
        batch = min(spec['batch'], 512 if name == 'cifar' else 16)
        steps = max(3, min(steps_for(args, name), 30 if name == 'cifar' else 4))
        base = cpu_reference_run(spec, steps, max(min(args.warmup, 3 if name == 'cifar' else 1), 1), seconds=60.0 if name == 'cifar' else 45.0, batch=batch)
        results[name] = dict(value=base['value'], unit=UNIT, ms_per_step=base['ms_per_step'], per_step_batch=batch, config=dict(workload=spec['label']),
                             cpu_baseline=dict(value=base['value'], unit=UNIT, cores=base['cores'], kind=base['kind'], sample=base['sample']))
    head = results[names[0]]
    line = dict(metric=METRIC, value=head['value'], unit=UNIT, n_gpus=args.gpus, steps=steps_for(args, names[0]), warmup=args.warmup, ms_per_step=head['ms_per_step'], higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=head['config']['workload'], per_step_batch=head['per_step_batch'],
                            note='reference CPU path = oracle restatement (the reference package cannot be imported here: SURVEY.md section 8.c)'),
                cpu_baseline=head['cpu_baseline'], e2e=dict(value=head['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0, workloads=results)
    print(json.dumps(line))


def _graph_time_ms(launch, reps, iters, stream):
    """ ms per launch: `reps` launches (distinct buffers, working set > L2) captured into one CUDA graph on `stream`, replayed `iters` times between CUDA
    events recorded on that same stream. Eager Python + ctypes launches take longer than these kernels run, so they would time the host. """
    import torch
    with torch.cuda.stream(stream):
        for i in range(reps):
            launch(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        for i in range(reps):
            launch(i)
    graph.replay()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        start.record()
        for _ in range(iters):
            graph.replay()
        end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / (iters * reps)


C4_LAYERS = [  # name, c, h, k, ksize, launches per step (forward, data gradient, weight gradient) in conf/base/resnet_style.yml
    ('s1 64->64 @56', 64, 56, 64, 3, (4, 4, 4)), ('s2in 64->128 @28', 64, 28, 128, 3, (1, 1, 1)), ('s2 128->128 @28', 128, 28, 128, 3, (4, 4, 4)),
    ('s3in 128->256 @14', 128, 14, 256, 3, (1, 1, 1)), ('s3 256->256 @14', 256, 14, 256, 3, (4, 4, 4)), ('s4in 256->512 @7', 256, 7, 512, 3, (1, 1, 1)),
    ('s4 512->512 @7', 512, 7, 512, 3, (2, 2, 2)),
]


def conv_layer_table(batch, dev, peaks):
    """ Forward / data-gradient / weight-gradient convolution of every ResNet-style layer shape through the C ABI, timed live (graph replay, buffers cycled
    beyond L2): TFLOP/s on 2*N*P*Q*K*C*R*S FLOP and the fraction of the measured dense bf16 peak = the "conv tensor-pipe %" of BASELINE.json's metric. """
    import ctypes
    import torch
    from deepcv_b200._lib import ACT_LEAKY_RELU, ALGO_AUTO, DCV_BF16, ConvShape, check, lib
    stream = torch.cuda.Stream()
    st = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    rows, tot_flop, tot_ms = [], 0., 0.
    with torch.cuda.stream(stream):
        for name, c, h, k, ks, counts in C4_LAYERS:
            n, pad = batch, ks // 2
            shape = ConvShape(n, h, h, c, k, ks, ks, 1, 1, pad, pad, 1, 1, h, h)
            reps = max(2, int(300e6 // (n * h * h * (c + k) * 2)) + 1)
            xs = [torch.randn(n, h, h, c, device=dev).bfloat16() for _ in range(reps)]
            ys = [torch.randn(n, h, h, k, device=dev).bfloat16() for _ in range(reps)]
            w = (torch.randn(k, ks, ks, c, device=dev) * 0.05).bfloat16()
            wt = torch.empty(c, ks, ks, k, device=dev, dtype=torch.bfloat16)
            w32 = w.float()
            check(lib.dcv_pack_conv_weight(P(w32), P(wt), DCV_BF16, k, ks, ks, c, 1, st), 'pack')
            bias, dw = torch.zeros(k, device=dev), torch.empty(k, ks, ks, c, device=dev)
            flop = 2.0 * n * h * h * k * c * ks * ks
            row = dict(layer=name)
            ops_ = dict(fwd=lambda i: check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xs[i]), P(w), P(bias), P(ys[i]), None, ACT_LEAKY_RELU, 0.01, DCV_BF16, ALGO_AUTO, 0, st), 'fwd'),
                        dgrad=lambda i: check(lib.dcv_conv2d_dgrad(ctypes.byref(shape), P(ys[i]), P(w), P(wt), P(xs[i]), DCV_BF16, ALGO_AUTO, st), 'dgrad'),
                        wgrad=lambda i: check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xs[i]), P(ys[i]), P(dw), None, DCV_BF16, ALGO_AUTO, 0, st), 'wgrad'))
            for j, (op, fn) in enumerate(ops_.items()):
                ms = _graph_time_ms(fn, reps, 3, stream)
                row[op] = dict(us=round(ms * 1e3, 1), tflops=round(flop / ms / 1e9, 1), frac=round(flop / ms / 1e9 / peaks['bf16_tflops'], 3))
                tot_flop += flop * counts[j]; tot_ms += ms * counts[j]
            rows.append(row)
            del xs, ys
    return dict(per_layer=rows, weighted_conv_tflops=tot_flop / tot_ms / 1e9, weighted_conv_frac=tot_flop / tot_ms / 1e9 / peaks['bf16_tflops'],
                note='3x3 layers of the step weighted by their launch counts; the 7x7/2 stem (gather kernels) is not in this table')


def time_dominant_kernel(workload, batch, size, dev, peaks, dtype_name):
    """ Roofline of the dominant kernel of the step, called through the C ABI with preallocated buffers, timed live with CUDA events on the stream it is
    launched on (graph replay, buffers cycled beyond L2).
      cifar     the fused few-channel backward kernel (data + weight gradient in one launch) of the 5x5 4 -> 4 channel layers at 32x32: the largest share
                of the CIFAR step in the ncu launch list (profiles/r02_launch_summary_cifar_warm.txt). HBM-bound by the accounting of SURVEY.md section
                8.d (arithmetic intensity 43-72 FLOP/B): algorithmic bytes per launch = read x + read dz + read y + write dx + write dw.
      imagenet  the tcgen05 implicit-GEMM forward convolution of the 64 -> 64 channel 3x3 layers at 56x56 (also run as their data gradient): the largest
                share of the ImageNet-shaped step. Tensor-bound: 2*N*P*Q*K*C*R*S FLOP per launch against the measured dense bf16 peak. """
    import ctypes
    import torch
    from deepcv_b200._lib import ACT_LEAKY_RELU, ALGO_AUTO, ALGO_DIRECT, DCV_BF16, DCV_F32, ConvShape, check, lib
    tdt, dt, esize = (torch.bfloat16, DCV_BF16, 2) if dtype_name == 'bf16' else (torch.float32, DCV_F32, 4)
    stream = torch.cuda.Stream()
    st = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.stream(stream):
        if workload == 'cifar' and dtype_name == 'bf16':
            # the fused few-channel backward kernel of the 4 -> 4 channel 5x5 layers (2 launches per step + the same kernel at 16x16; the largest share of the
            # launch list): reads the layer input x, the incoming gradient dz and the layer's own output y (the activation derivative is applied while
            # loading), writes the data gradient dx and accumulates dw / dbias
            from deepcv_b200._lib import ACT_RELU
            n, c, h, w, k = batch, 4, size, size, 4
            reps = max(4, int(300e6 / (n * h * w * (2 * c + 2 * k) * esize)) + 1)
            xs = [torch.randn(n, h, w, c, device=dev).to(tdt) for _ in range(reps)]
            dzs = [torch.randn(n, h, w, k, device=dev).to(tdt) for _ in range(reps)]
            ys = [torch.randn(n, h, w, k, device=dev).relu().to(tdt) for _ in range(reps)]
            dxs = [torch.empty(n, h, w, c, device=dev, dtype=tdt) for _ in range(reps)]
            wt = (torch.randn(k, 5, 5, c, device=dev) * 0.1).to(tdt)
            dw, db = torch.zeros(k, 5, 5, c, device=dev), torch.zeros(k, device=dev)
            shape = ConvShape(n, h, w, c, k, 5, 5, 1, 1, 2, 2, 1, 1, h, w)
            assert lib.dcv_sc_conv_bwd_supported(ctypes.byref(shape), DCV_BF16), 'the fused few-channel backward kernel does not serve the 4 -> 4 channel 5x5 layer'

            def launch(i):
                check(lib.dcv_sc_conv_bwd(ctypes.byref(shape), P(xs[i]), None, P(dzs[i]), P(ys[i]), None, ACT_RELU, 0., P(wt), P(dxs[i]), P(dw), P(db), None, None, None, None, st), 'sc_conv_bwd')
            ms = _graph_time_ms(launch, reps, 20, stream)
            alg_bytes = n * h * w * (2 * c + 2 * k) * esize + k * 25 * c * 4
            achieved = alg_bytes / (ms / 1e3) / 1e9
            traffic = _committed_traffic('sc_bwd_kernel<4, 4, 5>') if n == 512 else None
            return dict(bound='hbm', kernel='sc_bwd_kernel<4, 4, 5> (4->4 ch, 5x5, 32x32: data gradient + weight gradient in one launch on mma.sync, activation derivative applied on load; '
                        'timed without the pending-normalisation terms the step adds)', achieved=achieved,
                        peak=peaks['hbm_gbs'], unit='GB/s', frac=achieved / peaks['hbm_gbs'], traffic=traffic, traffic_unit='bytes of DRAM per launch',
                        traffic_source='ncu --set full (cold caches), profiles/r02_traffic.json' if traffic else None,
                        peak_source=peaks['source'], algorithmic_bytes_per_launch=alg_bytes, us_per_launch=ms * 1e3,
                        note='latency-bound at this size: 16.8 MB per launch is 2.5 us of HBM time; the kernel is a chain of dependent stage -> MMA -> reduce phases per image pair (see DESIGN.md section 5)')
        if workload == 'cifar':   # fp32 parity mode: the CUDA-core direct weight gradient
            n, c, h, w, k = batch, 4, size, size, 4
            reps = max(4, int(300e6 / (n * h * w * (c + k) * esize)) + 1)
            xs = [torch.randn(n, h, w, c, device=dev).to(tdt) for _ in range(reps)]
            dys = [torch.randn(n, h, w, k, device=dev).to(tdt) for _ in range(reps)]
            dw = torch.empty(k, 5, 5, c, device=dev)
            shape = ConvShape(n, h, w, c, k, 5, 5, 1, 1, 2, 2, 1, 1, h, w)

            def launch(i):
                check(lib.dcv_conv2d_wgrad(ctypes.byref(shape), P(xs[i]), P(dys[i]), P(dw), None, dt, ALGO_DIRECT, 0, st), 'conv2d_wgrad')
            ms = _graph_time_ms(launch, reps, 20, stream)
            alg_bytes = n * h * w * (c + k) * esize + k * 25 * c * 4
            achieved = alg_bytes / (ms / 1e3) / 1e9
            return dict(bound='hbm', kernel='conv_wgrad_direct_s1_kernel<5,32,16,4,4> (4->4 ch, 5x5, 32x32; includes the 1.6 KB memset of dw)', achieved=achieved, peak=peaks['hbm_gbs'], unit='GB/s',
                        frac=achieved / peaks['hbm_gbs'], traffic=None, peak_source=peaks['source'], algorithmic_bytes_per_launch=alg_bytes, us_per_launch=ms * 1e3)
        n, c, h, w, k = batch, 64, 56, 56, 64
        if dtype_name != 'bf16':
            return None
        reps = max(3, int(400e6 / (n * h * w * (c + k) * esize)) + 1)
        xs = [torch.randn(n, h, w, c, device=dev).to(tdt) for _ in range(reps)]
        ys = [torch.empty(n, h, w, k, device=dev, dtype=tdt) for _ in range(reps)]
        wt = (torch.randn(k, 3, 3, c, device=dev) * 0.05).to(tdt)
        bias = torch.zeros(k, device=dev)
        shape = ConvShape(n, h, w, c, k, 3, 3, 1, 1, 1, 1, 1, 1, h, w)

        def launch(i):
            check(lib.dcv_conv2d_fwd(ctypes.byref(shape), P(xs[i]), P(wt), P(bias), P(ys[i]), None, ACT_LEAKY_RELU, 0.01, dt, ALGO_AUTO, 0, st), 'conv2d_fwd')
        ms = _graph_time_ms(launch, reps, 10, stream)
        flop = 2.0 * n * h * w * k * c * 9
        achieved = flop / (ms / 1e3) / 1e12
        # traffic: DRAM bytes of one `ncu --set full` capture of this kernel at batch 256 (profiles/r01_ncu_full_summary.txt: 102.9 MB read = the input tensor once,
        # 50.7 MB written before the kernel ends — the rest of the 102.8 MB output is still in L2)
        traffic = 153.6e6 if n == 256 else None
        return dict(bound='tensor', kernel='conv_fwd_tc_halo_kernel<64> (64->64 ch, 3x3, 56x56, bias + LeakyReLU epilogue)', achieved=achieved, peak=peaks['bf16_tflops'], unit='TFLOP/s',
                    frac=achieved / peaks['bf16_tflops'], traffic=traffic, traffic_unit='bytes of DRAM per launch', traffic_source='ncu --set full, profiles/r01_ncu_full_summary.txt',
                    peak_source=peaks['source'] + ', burst figure (kernel timed alone)', algorithmic_flop_per_launch=flop, us_per_launch=ms * 1e3)


def _committed_traffic(kernel_name):
    """ DRAM bytes (read + written) of one launch of `kernel_name` from the committed `ncu --set full` capture digest (profiles/r02_traffic.json,
    written by tools/ncu_traffic.py from `ncu -i X.ncu-rep --page raw --csv`), or None. """
    p = ROOT / 'profiles' / 'r02_traffic.json'
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(kernel_name, {}).get('dram_bytes')


def load_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d['hbm_gbs']), bf16_tflops=float(d['bf16_tflops']), bf16_tflops_sustained=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


TRAIN_FLOP_PER_IMAGE = dict(cifar=10.64e6, imagenet=11.23e9)   # convolution FLOP of one training step per image
HBM_FLOOR_BYTES_PER_IMAGE = dict(cifar=0.38e6)                # ideal-fusion HBM traffic of the default net per image


def measure_workload(name, args, dev, world, rank, local_rank, peaks):
    """ One workload end to end: builds the model behind the public API, captures the step, times `value` (device-resident pool), `e2e` (pinned host
    batches through Engine.run) and the roofline of its dominant kernel. Collective (barriers): every rank calls it. """
    import gc
    from collections import OrderedDict
    import torch
    import torch.distributed as dist
    from deepcv_b200 import ops
    from deepcv_b200.meta.base_module import DeepcvModule
    from deepcv_b200.meta.data.datasets import dataloader_prefetch_batches
    from deepcv_b200.meta.data.preprocess import FusedPreprocess
    from deepcv_b200.meta.flat_params import FlatAdamW, flatten_parameters
    from deepcv_b200.meta.ignite_training import CrossEntropyLoss, DataParallelModel, Engine, make_process_function

    spec = workload_spec(name, args.batch)
    batch, size, classes = spec['batch'], spec['size'], spec['classes']
    steps, warmup = steps_for(args, name), max(args.warmup, 3)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32

    torch.manual_seed(563454)   # same initial weights on every rank (DDP broadcasts rank 0's; same seed + broadcast in DataParallelModel)
    model = DeepcvModule((3, size, size), spec['hp']).to(dev)
    pre = FusedPreprocess(mean=spec['mean'], std=spec['std'], pad=spec['pad'], flip=True, dtype=dtype, seed=434546 + rank).to(dev)
    net = DataParallelModel(model) if world > 1 else model
    flat = flatten_parameters(model)
    opt = FlatAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1. / world).attach(flat)
    if world > 1:
        net.reducer.average_in_finish = False   # 1 / world is applied inside the AdamW kernel
    losses = OrderedDict(main_loss=CrossEntropyLoss())

    # synthetic data: a pool of distinct uint8 batches larger than L2 (126 MB), cycled
    g = torch.Generator().manual_seed(563454 + rank)
    batch_bytes = batch * size * size * 3
    pool_n = max(4, min(256, int(200e6 // batch_bytes) + 1))
    pool = [torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8) for _ in range(pool_n)]
    labels = [torch.randint(0, classes, (batch,), generator=g) for _ in range(pool_n)]
    pool_dev = [p.to(dev) for p in pool]
    labels_dev = [l.to(dev) for l in labels]
    n_host = min(8, pool_n)
    pool_host = [p.pin_memory() for p in pool[:n_host]]
    labels_host = [l.pin_memory() for l in labels[:n_host]]
    del pool, labels

    # the public training step: what train() hands to the Engine. Its first call captures the CUDA graph (and rewinds the model / optimizer state).
    launches0 = ops.launch_count()
    step_fn = make_process_function({}, dev, net, losses, opt, preprocess=pre, cuda_graph=False if args.no_graph else None)
    trainer = Engine(step_fn)
    trainer.run([(pool_host[0], labels_host[0])], max_epochs=1)
    if args.no_graph:
        launches_per_step = ops.launch_count() - launches0
        runner = None
    else:
        runner = next(iter(step_fn.runners.values()))
        launches_per_step = (ops.launch_count() - launches0) // 4   # 3 warm-up steps + 1 captured step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    window = {}

    def timed(fn, n_warm):
        fn(n_warm, warm=True)
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        window['t0'] = time.time()
        start.record()
        fn(None, warm=False)
        end.record()
        barrier()
        window['t1'] = time.time()
        ms = start.elapsed_time(end)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident throughput: the captured step replayed over batches already in HBM
    def resident(n_warm, warm):
        for i in range(n_warm if warm else steps):
            if runner is not None:
                runner.step(pool_dev[i % pool_n], labels_dev[i % pool_n])
            else:
                step_fn(trainer, (pool_dev[i % pool_n], labels_dev[i % pool_n]))
    remeasured = False
    with ClockSampler(local_rank) as clocks:
        clocks.wait_ready()
        ms = timed(resident, warmup)
        clock_window = (window['t0'], window['t1'])
        again = needs_remeasure(clocks.summary(*clock_window))
        if world > 1:   # collective decision: `timed` contains barriers
            flag = torch.tensor([int(again)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            again = bool(int(flag.item()))
        if again:
            # hardware / thermal slowdown, or clocks far below max with no reason given (a leftover lock): the number is rejected and taken ONCE more
            time.sleep(2.0)
            remeasured = True
            ms = timed(resident, warmup)
            clock_window = (window['t0'], window['t1'])
    value = world * batch * steps / (ms / 1e3)

    # ---- end to end through the public API: Engine.run over a loader of pinned host uint8 batches; per step H2D of images / labels / augmentation
    # parameters (+ learning rate when it changes), the graph replay, and the D2H read of the loss (`process_function` returns floats)
    e2e_steps = max(10, steps)   # as many steps as the device-resident measurement: a single host hiccup (one slow iteration) weighs less
    losses_seen = []
    trainer.add_event_handler(__import__('deepcv_b200.meta.ignite_training', fromlist=['Events']).Events.ITERATION_COMPLETED, lambda e: losses_seen.append(e.state.output['main_loss']))

    def e2e_run(n_warm, warm):
        n = n_warm if warm else e2e_steps
        # `dataloader_prefetch_batches` (reference meta/data/datasets.py:76-115, what train() applies to a pinned loader): batch i + 1 is copied while step i runs
        trainer.run(dataloader_prefetch_batches([(pool_host[i % n_host], labels_host[i % n_host]) for i in range(n)], dev), max_epochs=trainer.state.epoch + 1)
    gc.collect()   # garbage of an earlier workload (its captured graph, pinned pools) must not be collected — cudaFree synchronises — inside the timed loop
    torch.cuda.synchronize()
    ms_e2e = timed(e2e_run, 3)
    e2e_value = world * batch * e2e_steps / (ms_e2e / 1e3)
    if os.environ.get('DCV_BENCH_PROFILE_E2E'):   # tuning aid: where the HOST time of the end-to-end loop goes (cProfile, printed to stderr, not timed)
        import cProfile, pstats
        prof = cProfile.Profile()
        prof.enable()
        e2e_run(None, warm=False)
        torch.cuda.synchronize()
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats('tottime').print_stats(22)
    h2d = batch_bytes + batch * 8 + batch * (1 + 8) + 4    # images + int64 labels + flip (u8) / crop (2 x i32) + learning rate
    e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4 * len(losses), ms_per_step=ms_e2e / e2e_steps, steps=e2e_steps,
               api='ignite_training.Engine(make_process_function(...)).run(dataloader_prefetch_batches(loader of pinned uint8 batches))')

    result = dict(value=value, unit=UNIT, steps=steps, warmup=warmup, ms_per_step=ms / steps, e2e=e2e, gpu_launches=int(launches_per_step * steps), launches_per_step=int(launches_per_step),
                  clocks=dict(clocks.summary(*clock_window), remeasured=remeasured),
                  config=dict(workload=spec['label'], per_gpu_batch=batch, global_batch=batch * world, parallelism=f'dp{world}',
                              step='fused u8 preprocess(normalise+flip+crop) + fwd + CE + bwd + ' + ('bucketed NCCL all-reduce + ' if world > 1 else '') + 'AdamW'
                                   + ('' if args.no_graph else ', one CUDA graph replay per step'),
                              l2='inputs cycle through a pool of %d distinct uint8 batches (%.0f MB > 126 MB L2)' % (pool_n, pool_n * batch_bytes / 1e6),
                              final_loss=losses_seen[-1] if losses_seen else None))
    flop = TRAIN_FLOP_PER_IMAGE[name] * batch
    result['step_conv_tflops'] = flop / (ms / steps) / 1e9
    if name == 'imagenet':
        result['step_frac_of_sustained_bf16_peak'] = result['step_conv_tflops'] / peaks['bf16_tflops_sustained']
    else:
        result['step_frac_of_hbm_floor'] = (HBM_FLOOR_BYTES_PER_IMAGE[name] * batch / (ms / steps) / 1e6) / peaks['hbm_gbs']

    # release the captured graph and the model before the next workload
    if world > 1 and runner is not None:
        runner.graph.reset()
    del step_fn, trainer, runner, net, model, opt, flat, pool_dev, labels_dev, pool_host, labels_host
    gc.collect()
    torch.cuda.empty_cache()

    if rank == 0:
        roofline = time_dominant_kernel(name, batch, size, dev, peaks, args.dtype)
        if name == 'imagenet' and roofline is not None and not args.no_layer_table:
            roofline['conv_layers'] = conv_layer_table(batch, dev, peaks)
        result['roofline'] = roofline
        result['cpu_baseline'] = None
        if not args.no_cpu_baseline and world == 1:
            result['cpu_baseline'] = cpu_reference_run(spec, steps=30 if name == 'cifar' else 3, warmup=2 if name == 'cifar' else 1, seconds=args.cpu_seconds if name == 'cifar' else args.cpu_seconds / 2,
                                                       batch=min(batch, 512 if name == 'cifar' else 8))
        torch.cuda.empty_cache()
    barrier()
    return result


def run_b200(args):
    import torch
    import torch.distributed as dist
    from deepcv_b200._lib import check, lib

    rank, world, local_rank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    check(lib.dcv_device_check(), 'device_check')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()
    names = ['cifar', 'imagenet'] if args.workload == 'all' else [args.workload]
    results = {}
    for name in names:
        results[name] = measure_workload(name, args, dev, world, rank, local_rank, peaks)
        gc.collect()   # the workload's captured graph, pools and pinned batches go NOW (not in the middle of the next workload's timed loop)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def shutdown():
        # A captured CUDA graph that contains NCCL kernels keeps the communicator busy: destroy_process_group() was seen to hang on it.
        # The graphs were released per workload; leave the process without the collective teardown.
        if world > 1:
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        shutdown()
        return
    head = results[names[0]]
    line = dict(metric=METRIC, value=head['value'], unit=UNIT, n_gpus=world, steps=head['steps'], warmup=head['warmup'], ms_per_step=head['ms_per_step'], higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype=args.dtype, data='synthetic', config=head['config'], e2e=head['e2e'], gpu_launches=sum(r['gpu_launches'] for r in results.values()),
                launches_per_step=head['launches_per_step'], clocks=head['clocks'], roofline=head.get('roofline'), cpu_baseline=head.get('cpu_baseline'), workloads=results)
    print(json.dumps(line))
    shutdown()


def run_preprocess_sweep(args):
    """ BASELINE.json configs[4]: the uint8 normalise / flip / crop kernel alone, 32^2 .. 1024^2, achieved HBM GB/s on algorithmic bytes
    (3*h*w*(1 + sizeof(out)) per image) against the measured copy bandwidth. Runs tools/elementwise_bench.py and folds its lines into one. """
    if int(os.environ.get('RANK', 0)) != 0:
        return
    out = subprocess.run([sys.executable, str(ROOT / 'tools' / 'elementwise_bench.py'), '--what', 'preprocess'], capture_output=True, text=True, check=True).stdout
    rows = [json.loads(l) for l in out.splitlines() if l.startswith('{')]
    peaks = load_peaks()
    best = max(rows, key=lambda r: r['GBps'])
    print(json.dumps(dict(metric='preprocess achieved HBM bandwidth', value=best['GBps'], unit='GB/s', n_gpus=1, higher_is_better=True, dtype='u8', data='synthetic',
                          config=dict(workload='uint8 HWC -> pad-crop -> flip -> normalise sweep 32^2..1024^2 (bf16 and fp32 outputs), ~1 GB of traffic per launch'),
                          roofline=dict(bound='hbm', achieved=best['GBps'], peak=peaks['hbm_gbs'], unit='GB/s', frac=best['GBps'] / peaks['hbm_gbs'], traffic=None, kernel=best['kernel']),
                          sweep=rows)))


def main():
    args = parse_args()
    if os.environ.get('DCV_BENCH_WATCHDOG'):   # debugging aid: dump all Python stacks and exit if the run is still going after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['DCV_BENCH_WATCHDOG']), exit=True)
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'preprocess':
        run_preprocess_sweep(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
