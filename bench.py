#!/usr/bin/env python
"""bench.py — DM + S2D distillation iterations/sec on synthetic miniUCF101-shaped data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], flags of sh/s2d/s2d_DM_ms.sh): distill_s2d_ms.py --method DM,
50 classes, videos 16x3x112x112, vpc=1 spc=2 dpc=2, batch_real=64, --no_train_static.  One "step" =
one full DM iteration: fresh random frozen ConvNet3D, composer, 50x64 real + 50 synthetic video
embeddings, DM loss, backward to dynamic memory + hallucinator, momentum-SGD updates.

* precision (default f16x3): the parity mode of the fused tcgen05 pipeline — fp16 hi/lo operand pairs, three products per
          MAC into one fp32 TMEM accumulator (embeddings ~5e-5 of fp32, routing-conditioned gradients < 1e-3; tests/test_x3_gpu.py,
          tests/test_parity_fullsize_gpu.py).  `throughput_mode` reports the single-pass bf16 line beside it.
* value : iterations/s with the real set resident in HBM (device-timed, max over ranks).
* e2e   : same iteration driven from HOST memory the way the reference holds its data (get_images, distill_s2d_ms.py:81-87):
          every step copies its sampled real videos as normalised fp32 from pinned host memory (double-buffered on a copy
          stream) and reads the loss back.  `e2e_uint8_host` (decoded uint8 frames on the host, normalisation fused into the
          packer, a quarter of the PCIe bytes) and `e2e_resident` (dataset uploaded once) are reported next to it.
* check : loss of the last timed step and checksums of the trained memories after the timed steps — equal for every --gpus N.
* roofline : conv-1 tcgen05 kernel (69 % of the FLOPs), CUDA events around its launches, algorithmic FLOPs.
* reference_torch_cuda : the reference's OWN modules (baseline/_ref) running the verbatim loop body on this GPU (cuDNN),
          allow_tf32 off and on (BASELINE.json configs[1] "vs reference torch-CUDA").
* cpu_baseline : the reference's own modules on the host cores, bounded sample (kind "reference"; the oracle port if absent).
With --impl reference the reference's CPU path alone is timed (rank 0 only) on the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, T, HW, VPC, SPC, DPC, BATCH_REAL, PER_CLASS = 50, 16, 112, 1, 2, 2, 64, 72
F_L0, F_L1, F_L2 = 2.832e9, 7.553e9, 0.617e9          # algorithmic FLOP per video
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]     # dataset normalisation (utils.py:214-230)
F_EMBED = F_L0 + F_L1 + F_L2
# learning rates of both arms (script arguments L_D / L_H of the reference's sh/s2d/s2d_DM_ms.sh): small enough that the memories stay
# finite on random data over any number of bench iterations, so that `check` is a meaningful fingerprint
LR_DYNAMIC, LR_HAL = 1.0, 1e-4


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.window = None           # (t0, t1) host times of the timed region: only samples inside it are reported

    def mark(self, t0, t1):
        self.window = (t0, t1)

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.idx}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')] + [time.time()])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        if self.window is not None:
            inside = [r for r in self.rows if self.window[0] <= r[-1] <= self.window[1] + 0.05]
            if inside:
                self.rows = inside
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 9 and r[5 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max([float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace('.', '').isdigit()] or [0.0]),
                'samples': len(sm), 'reasons': reasons}


# ------------------------------------------------------------------------------------------ reference arm
def reference_dir():
    """Where the reference's own python modules live: baseline/_ref (staged from /root/reference by __graft_entry__.build();
    git-ignored, travels to the GPU box) -> $VD_REFERENCE -> /root/reference.  None when absent."""
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), os.environ.get('VD_REFERENCE'), '/root/reference'):
        if cand and os.path.exists(os.path.join(cand, 'utils.py')) and os.path.exists(os.path.join(cand, 'networks.py')):
            return cand
    return None


_REF_MODS = {}


def import_reference(path):
    """The reference's top-level modules (utils, networks) imported from `path` without being shadowed by this repo's
    drop-in modules of the same names."""
    if path in _REF_MODS:
        return _REF_MODS[path]
    saved = list(sys.path)
    hidden = {n: sys.modules.pop(n) for n in ('networks', 'utils', 'reparam_module', 'distill_utils') if n in sys.modules}
    sys.path[:] = [path] + [q for q in saved if os.path.abspath(q or '.') != ROOT]
    try:
        import networks as ref_networks          # noqa
        import utils as ref_utils                # noqa
    finally:
        for n in ('networks', 'utils', 'reparam_module', 'distill_utils', 'distill_utils.dataset'):
            sys.modules.pop(n, None)
        sys.modules.update(hidden)
        sys.path[:] = saved
    _REF_MODS[path] = (ref_utils, ref_networks)
    return _REF_MODS[path]


class ReferenceLoop:
    """State of distill_s2d_ms.py:89-108 and the VERBATIM iteration body of :393-438 (--method DM, --no_train_static) driven
    through the reference's own modules (utils.get_network -> networks.ConvNet3D, utils.Conv3DNet, torch.optim.SGD) on
    `device`.  The real set is a host TensorDataset of normalised fp32 videos like the reference's --preload.  `n_cls` < C
    runs a bounded sample of the workload: the first n_cls classes of the per-class loop (the loop body is per class)."""

    def __init__(self, ref_utils, device, n_cls, videos_host=None, per_class=None):
        self.u, self.device, self.n_cls = ref_utils, device, n_cls
        per_class = per_class or (BATCH_REAL + 2)
        if videos_host is None:
            g = torch.Generator().manual_seed(0)
            videos_host = torch.randn(n_cls * per_class, T, 3, HW, HW, generator=g)
        labels = torch.arange(n_cls).repeat_interleave(per_class)
        self.dst_train = torch.utils.data.TensorDataset(videos_host, labels)
        self.indices_class = [list(range(c * per_class, (c + 1) * per_class)) for c in range(n_cls)]
        torch.manual_seed(0)
        self.static_syn = torch.randn(size=(C * SPC, 3, HW, HW), dtype=torch.float).detach().to(device).requires_grad_(False)
        self.dynamic_syn = torch.randn(size=(C, DPC, T, 1, HW, HW), dtype=torch.float).detach().to(device).requires_grad_(True)
        self.hals = torch.nn.ModuleList([ref_utils.Conv3DNet()]).to(device)
        self.optimizer_dynamic = torch.optim.SGD([self.dynamic_syn], lr=LR_DYNAMIC, momentum=0.95)
        self.optimizer_hals = torch.optim.SGD(self.hals.parameters(), lr=LR_HAL, momentum=0.95)

    def get_images(self, c, n):                                                      # distill_s2d_ms.py:81-87
        idx_shuffle = np.random.permutation(self.indices_class[c])[:n]
        imgs = torch.cat([self.dst_train[i][0].unsqueeze(0) for i in idx_shuffle], 0)
        return imgs.to(self.device)

    def iteration(self):
        dev, num_classes, vpc, spc = self.device, C, VPC, SPC
        net = self.u.get_network('ConvNet3D', 3, num_classes, (HW, HW), frames=T, dist=False).to(dev)  # get a random model (:393)
        net.train()
        for param in list(net.parameters()):
            param.requires_grad = False
        embed = net.embed
        label = torch.tensor(np.stack([np.ones(vpc) * i for i in range(0, num_classes)]), dtype=torch.long, requires_grad=False, device=dev).view(-1)
        ran = torch.arange(0, num_classes * vpc).to(dev)
        idx = ran % vpc
        dynamic_idx = 2 * idx + torch.randint(2, (num_classes * vpc,), device=dev)
        static_idx = spc * label + 2 * idx + torch.randint(2, (num_classes * vpc,), device=dev)
        static = self.static_syn[static_idx, :, :, :]
        dynamic = self.dynamic_syn[label, dynamic_idx, :, :, :, :]
        hal = self.hals[0]
        image_syn = hal(static, dynamic)
        loss = torch.tensor(0.0).to(dev)
        for c in range(0, self.n_cls):                                               # the reference runs range(0, num_classes)
            img_real = self.get_images(c, BATCH_REAL)
            img_syn = image_syn[c * vpc:(c + 1) * vpc].reshape((vpc, T, 3, HW, HW))
            output_real = embed(img_real).detach()
            output_syn = embed(img_syn)
            loss += torch.sum((torch.mean(output_real, dim=0) - torch.mean(output_syn, dim=0)) ** 2)
        self.optimizer_dynamic.zero_grad()
        self.optimizer_hals.zero_grad()
        loss.backward()
        self.optimizer_dynamic.step()
        self.optimizer_hals.step()
        return loss.item()


def cpu_oracle_rate(n_classes, n_real, threads):
    """Fallback when the reference's modules are absent: one DM+S2D iteration of the CPU oracle port (oracle/dm.py) on
    n_classes classes; returns (it/s extrapolated to the full workload, seconds, description)."""
    import oracle
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    per = n_real + 2
    videos = torch.randn(n_classes * per, T, 3, HW, HW, generator=g)
    indices_class = [list(range(c * per, (c + 1) * per)) for c in range(n_classes)]
    static_syn = torch.randn(n_classes * SPC, 3, HW, HW, generator=g)
    dyn = torch.randn(n_classes, DPC, T, 1, HW, HW, generator=g)
    hal = oracle.init_hallucinator(1)
    params = oracle.init_convnet3d(2, 3, C)
    cd = torch.randint(2, (n_classes * VPC,), generator=g)
    cs = torch.randint(2, (n_classes * VPC,), generator=g)
    np.random.seed(0)
    t0 = time.perf_counter()
    r = oracle.dm_s2d_iteration(params, static_syn, dyn, hal, videos, indices_class, vpc=VPC, spc=SPC,
                                batch_real=n_real, coin_dynamic=cd, coin_static=cs)
    dt = time.perf_counter() - t0
    sample_units = n_classes * (n_real + 2 * VPC)
    full_units = C * (BATCH_REAL + 2 * VPC)
    its = 1.0 / (dt * full_units / sample_units)
    desc = (f'{n_classes} of {C} classes x ({n_real} real + {VPC} syn) videos {T}x3x{HW}x{HW}, one oracle-port DM+S2D iteration, '
            f'linearly extrapolated to {C} classes')
    assert torch.isfinite(r['loss'])
    return its, dt, desc


def reference_cpu_rate(loop, threads):
    """One iteration of the reference's own loop on the host cores over loop.n_cls classes -> (it/s extrapolated to all C
    classes, measured seconds, description).  The per-class loop body is the whole cost; the class-independent part
    (get_network, composer, optimiser) is timed inside and not scaled."""
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    loss = loop.iteration()
    dt = time.perf_counter() - t0
    assert np.isfinite(loss)
    its = 1.0 / (dt * C / loop.n_cls)
    desc = (f'{loop.n_cls} of {C} classes x ({BATCH_REAL} real + {VPC} syn) videos {T}x3x{HW}x{HW} through the reference\'s own '
            f'modules (utils.get_network / networks.ConvNet3D.embed / utils.Conv3DNet / torch.optim.SGD, loop body of '
            f'distill_s2d_ms.py:393-438), time scaled by {C}/{loop.n_cls}')
    return its, dt, desc


def bench_config(n_gpus):
    """The workload description shared verbatim by both arms (no implementation keys)."""
    return {'workload': WORKLOAD_DESC,
            'parallelism': (f'synthetic branch + memories sharded by class c%{n_gpus}, sampled real videos spread over the ranks, '
                            f'all-reduce of (C,D) partial embedding sums + [hal grad | loss] (NCCL)') if n_gpus > 1 else 'single GPU',
            'real_videos_per_step': C * BATCH_REAL, 'syn_videos_per_step': C * VPC,
            'l2_policy': f'inputs larger than L2: each step reads {C * BATCH_REAL} distinct real videos '
                         f'({C * BATCH_REAL * T * 3 * HW * HW * 4 / 1e9:.1f} GB fp32)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    total = max(1, args.warmup + args.steps)
    refdir = reference_dir()
    vals, secs = [], []
    if refdir is not None:
        ref_utils, _ = import_reference(refdir)
        # bounded sample: calibrate on 2 classes, then as many classes per step as fit ~150 s in total (at least 10)
        probe = ReferenceLoop(ref_utils, 'cpu', 2)
        _, dt2, _ = reference_cpu_rate(probe, threads)
        n_cls = int(max(min(10, C), min(C, (150.0 / total) / max(dt2 / 2, 1e-3))))
        np.random.seed(0)
        loop = ReferenceLoop(ref_utils, 'cpu', n_cls)
        for i in range(total):
            its, dt, desc = reference_cpu_rate(loop, threads)
            if i >= args.warmup:
                vals.append(its)
                secs.append(dt)
        kind, impl = 'reference', f'the reference\'s own modules from {os.path.relpath(refdir, ROOT) if refdir.startswith(ROOT) else refdir}, torch CPU, all host cores'
    else:
        n_cls = sample_classes(150.0 / total, threads)
        for i in range(total):
            its, dt, desc = cpu_oracle_rate(n_cls, BATCH_REAL, threads)
            if i >= args.warmup:
                vals.append(its)
                secs.append(dt)
        kind, impl = 'port', 'CPU oracle port (oracle/dm.py): reference modules not found'
    v = float(np.mean(vals))
    print(json.dumps({
        'impl': 'reference', 'metric': 'DM+S2D distill iters/sec', 'value': v, 'unit': 'it/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        # a step of this arm is the bounded sample: ms_per_step is what a step actually took, value is scaled to the workload
        'ms_per_step': 1000.0 * float(np.mean(secs)), 'ms_per_full_iteration': 1000.0 / v,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args.gpus), 'reference_impl': impl,
        'cpu_baseline': {'value': v, 'unit': 'it/s', 'cores': threads, 'kind': kind, 'sample': desc,
                         'sample_seconds': float(np.mean(secs))},
        'e2e': {'value': v, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def sample_classes(seconds, threads):
    """How many full classes of the workload fit in `seconds` of CPU-oracle-port time (calibrated on 2 classes x 8 videos)."""
    _, dt, _ = cpu_oracle_rate(2, 8, threads)
    per_class = dt / (2 * (8 + 2 * VPC)) * (BATCH_REAL + 2 * VPC)
    return int(max(1, min(C // 2, seconds / max(per_class, 1e-3))))


def reference_torch_cuda(dev, videos_host, iters=2):
    """The reference's own modules on this GPU (ATen -> cuDNN), verbatim loop body, all C classes, host-resident fp32 real set:
    it/s with cudnn.allow_tf32 off (fp32 parity setting) and on (PyTorch's default = the reference as shipped)."""
    refdir = reference_dir()
    if refdir is None:
        return {'unavailable': 'reference modules not found (baseline/_ref, $VD_REFERENCE, /root/reference)'}
    ref_utils, _ = import_reference(refdir)
    out = {'source': os.path.relpath(refdir, ROOT) if refdir.startswith(ROOT) else refdir, 'iterations_timed': iters,
           'note': 'reference loop body distill_s2d_ms.py:393-438 through utils.get_network / networks.ConvNet3D / utils.Conv3DNet on '
                   'cuda (cuDNN); real set = host TensorDataset of normalised fp32 videos, get_images(...).to(device) per class'}
    prev = torch.backends.cudnn.allow_tf32
    try:
        for name, tf32 in (('allow_tf32_false', False), ('allow_tf32_true', True)):
            torch.backends.cudnn.allow_tf32 = tf32
            np.random.seed(0)
            loop = ReferenceLoop(ref_utils, dev, C, videos_host=videos_host, per_class=PER_CLASS)
            loop.iteration()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                loss = loop.iteration()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / iters
            out[name] = {'value': 1.0 / dt, 'unit': 'it/s', 'ms_per_step': dt * 1e3, 'loss': loss}
            del loop
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return out


def memory_kernel_rooflines(step_fn, tr, steps=3):
    """Average device time of the memory-bound kernels of the step (CUPTI, sustained conditions) against
    their algorithmic bytes and the measured HBM copy bandwidth."""
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
    dur = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            d = dur.setdefault(e.name, [0, 0.0])
            d[0] += 1
            d[1] += (e.time_range.end - e.time_range.start) * 1e-3      # ms
    peaks, _ = measured_peaks()
    hbm = float(peaks['hbm_gbs'])
    n_syn = C * VPC
    thw, hw = T * HW * HW, HW * HW
    tcn = tr.embedder.tc
    p = tcn.plan
    dyp1 = getattr(tcn, 'dyp1_bytes_per_video', 0)
    # fp32 FMAs of the composer stencils (81 per output element forward; 81 + 81 backward): these kernels sit above the ridge
    # (5.5 FMA per byte) and are bound by the fp32 pipe — `fma_frac` is the fraction of the measured 36.0 TFMA/s
    # (scripts/microbench/ffma_rate.cu: 128 FMA / clk / SM at 1.965 GHz; FFMA2 issues two per instruction at the same rate)
    fma = {'compose_fwd_tma_kernel': n_syn * thw * 81, 'compose_fwd_tiled_kernel': n_syn * thw * 81,
           'compose_bwd_tma_kernel': n_syn * thw * 162, 'compose_bwd_fused_kernel': n_syn * thw * 162}
    rows = [  # (kernel-name substring, algorithmic bytes per launch, what)
        ('compose_fwd_tma_kernel', n_syn * 4 * (3 * hw + thw + 3 * thw), 'read static+dynamic (tensor-map TMA), write video'),
        ('compose_bwd_tma_kernel', n_syn * 4 * (3 * thw + thw + 3 * hw + thw), 'read d video + dynamic + static (tensor-map TMA), write d dynamic'),
        ('compose_fwd_tiled_kernel', n_syn * 4 * (3 * hw + thw + 3 * thw), 'read static+dynamic, write video'),
        ('compose_bwd_fused_kernel', n_syn * 4 * (3 * thw + thw + 3 * hw + thw), 'read d video + dynamic + static, write d dynamic'),
        ('compose_bwd_data_tiled_kernel', n_syn * 4 * (3 * thw + thw), 'read d video, write d dynamic'),
        ('compose_bwd_wdyn_tiled_kernel', n_syn * 4 * (3 * thw + thw), 'read d video + dynamic'),
        ('compose_bwd_weight_static_kernel', n_syn * 4 * (3 * thw + 3 * hw), 'read d video + static'),
        ('col2im_rows_kernel<7, 8', n_syn * (p.col0_bytes_per_video + 4 * 3 * thw), 'read conv-0 columns, write d video'),
        ('col2im_kernel<float', n_syn * (2 * p.col2_bytes_per_video + 4 * 128 * int(p.T2p) * int(p.H2p) * int(p.W2p)),
         'read fp32 conv-2 columns, write fp32 d input of conv 2 (split-bf16 backward, one of three passes)'),
        ('col2im_kernel<unsigned short', n_syn * (p.col2_bytes_per_video + 128 * int(p.T2p) * int(p.H2p) * int(p.W2p) + dyp1),
         'read conv-2 columns + codes, write padded planar dY1'),
        ('pack_video_kernel', n_syn * (4 * 3 * thw + p.x0_bytes_per_video), 'read fp32 video, write packed bf16 conv-0 operand'),
        ('pack_video_x3_kernel', n_syn * (4 * 3 * thw + getattr(tcn, 'x0_per', 0)), 'read fp32 video, write packed fp16 hi/lo conv-0 operand'),
        ('sgd_momentum_kernel', None, '20 B / element (dynamic memory)'),
        ('class_mean_kernel', C * BATCH_REAL * p.embed_dim * 4, 'read real embeddings'),
    ]
    out = []
    for key, nbytes, what in rows:
        hits = [(k, v) for k, v in dur.items() if key in k]
        if not hits:
            continue
        n = sum(v[0] for _, v in hits)
        ms_total = sum(v[1] for _, v in hits)
        if key == 'sgd_momentum_kernel':
            nbytes = 20 * tr.dynamic_syn.numel()
            per_step_ms = ms_total / steps                     # three launches per step; the dynamic-memory one dominates
            out.append({'kernel': key, 'ms': per_step_ms, 'bytes': nbytes, 'gbs': nbytes / per_step_ms / 1e6,
                        'frac': nbytes / per_step_ms / 1e6 / hbm, 'what': what})
            continue
        # the launches of one step together cover the step's synthetic videos `covers` times (batches above max_batch are cut into
        # several launches; the fp32 col2im runs once per pass of the split backward): bytes and time are both per step
        covers = 3 if key == 'col2im_kernel<float' else 1
        ms = ms_total / steps
        nbytes = covers * nbytes
        row = {'kernel': key, 'ms': ms / (n / steps), 'ms_per_step': ms, 'launches_per_step': n / steps, 'bytes': int(nbytes),
               'gbs': nbytes / ms / 1e6, 'frac': nbytes / ms / 1e6 / hbm, 'what': what}
        if key in fma:
            row.update({'bound': 'fp32 FMA', 'fma': int(fma[key]), 'tfma_s': fma[key] / ms / 1e9, 'fma_frac': fma[key] / ms / 1e9 / 36.0})      # per step, like the bytes
        else:
            row['bound'] = 'hbm'
        out.append(row)
    return {'peak_gbs': hbm, 'peak_kind': 'hbm_gbs of measured (copy read+write)', 'peak_tfma_s': 36.0, 'kernels': out}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import copy
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL's debug output (the version banner at NCCL_DEBUG=VERSION/WARN/INFO) goes to stdout by default, where rank 0 must
        # print exactly one JSON line: send it to stderr instead
        # (NCCL only honours NCCL_DEBUG_FILE above the VERSION level, so a preset NCCL_DEBUG=VERSION is raised to WARN)
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', ''):
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from video_distillation_b200 import _lib
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer, owned_classes

    # ---- synthetic real set: per-class generators so the data does not depend on the world size
    own = owned_classes(C, rank, world)
    labels = [c for c in range(C) for _ in range(PER_CLASS)]
    assert PER_CLASS % world == 0, 'the synthetic set spreads every class evenly over the ranks'
    per_rank = PER_CLASS // world                   # shard='video': positions p % world == rank of every class
    # frames are uint8 like decoded video, normalised with the dataset statistics (reference: utils.py:214-230); the fp32
    # tensor is what the reference's preloaded TensorDataset holds
    mean_t = torch.tensor(MEAN, device=dev).view(1, 1, 3, 1, 1)
    std_t = torch.tensor(STD, device=dev).view(1, 1, 3, 1, 1)
    c255 = torch.tensor(255.0, device=dev)          # tensor divisor: a Python scalar would become a reciprocal multiply
    frames = torch.empty(C * per_rank, T, 3, HW, HW, dtype=torch.uint8, device=dev)
    vids = torch.empty(C * per_rank, T, 3, HW, HW, device=dev)
    for c in range(C):
        g = torch.Generator(device=dev).manual_seed(1000 + c)
        fr = torch.randint(0, 256, (PER_CLASS, T, 3, HW, HW), dtype=torch.uint8, device=dev, generator=g)[rank::world]
        frames[c * per_rank:(c + 1) * per_rank] = fr
        vids[c * per_rank:(c + 1) * per_rank] = ((fr.float() / c255) - mean_t) / std_t       # IEEE divisions, like the host transform
        del fr
    frames_host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
    frames_host.copy_(frames)
    del frames
    ds = DeviceDataset.from_device_shard(vids, labels, C, dev, rank, world, shard='video')

    def make_trainer(precision, dataset):
        torch.manual_seed(0)
        return DMS2DTrainer(dataset, num_classes=C, im_size=(HW, HW), frames=T, vpc=VPC, spc=SPC, dpc=DPC, batch_real=BATCH_REAL,
                            lr_dynamic=LR_DYNAMIC, lr_hal=LR_HAL, precision=precision, device=dev, init_on_device=True,
                            max_batch=args.max_batch,
                            syn_on_tensor_cores={'fused': True, 'split': 'split', 'fp32': False}[args.syn_mode])
    tr = make_trainer(args.precision, ds)
    prepack = None
    if tr.embedder.tc is not None and not args.no_prepack:
        # one-time conversion of the resident real set into the packed conv-0 operand (outside the timed region, like --preload)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ds.prepack(tr.embedder.tc, extra_slots=len(tr.owned) * tr.vpc)
        torch.cuda.synchronize()
        prepack = {'ms': (time.perf_counter() - t0) * 1e3, 'bytes_per_video': int(ds.x0.numel() // max(1, vids.shape[0] + (0 if tr.embedder.tc.real_products == 2 else len(tr.owned) * tr.vpc))),
                   'resident_bytes': int(ds.x0.numel()), 'videos': int(vids.shape[0])}
    np.random.seed(0)
    torch.cuda.manual_seed(1234)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    seed_box = [0]
    last_loss = [None]

    def step_resident(trainer=None):
        seed_box[0] += 1
        last_loss[0] = (trainer or tr).step(net_seed=seed_box[0])          # same seed on every rank -> same frozen net
        return last_loss[0]

    # ---- warm-up, then the timed region (device-resident inputs)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                         # started before the warm-up so that it is sampling when the timed region begins
    first = None
    for i in range(args.warmup):
        step_resident()
        if i == 0:
            # fingerprint of the FIRST iteration (identical state on every world size, no feedback yet): loss and the
            # dynamic-memory / hallucinator gradients must agree between --gpus 1, 2, 4, 8 to fp32 rounding
            gd = tr.dynamic_syn.grad.detach().double()
            part = torch.stack([gd.sum(), (gd * gd).sum()])
            if world > 1:
                dist.all_reduce(part)
            first = {'loss': float(last_loss[0]), 'grad_dynamic_sum': float(part[0]), 'grad_dynamic_sumsq': float(part[1]),
                     'grad_hal_weight_sumsq': float((tr.hal.encoder.weight.grad.double() ** 2).sum())}
            del gd
    tc = tr.embedder.tc
    if tc is not None:
        tc.timing = []
    _lib.launch_count_reset()
    t_wall0 = time.time()
    ms = timed(step_resident, args.steps)
    sampler.mark(t_wall0, time.time())
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    layer_ms = {0: 0.0, 1: 0.0, 2: 0.0}
    layer_videos = {0: 0, 1: 0, 2: 0}
    layer_launches = {0: 0, 1: 0, 2: 0}
    real_products = tc.real_products if tc is not None and tc.split else 1
    if tc is not None:
        # the launches of the frozen real videos (the differentiable synthetic videos of the two-product mode run in their own,
        # three-product launches: 1.5 % of the videos, counted in ms_per_step but not in the per-layer roofline figures)
        for layer, B, a, b, products in tc.timing:
            if products != real_products:
                continue
            layer_ms[layer] += a.elapsed_time(b)
            layer_videos[layer] += B
            layer_launches[layer] += 1
        tc.timing = None
    value = args.steps / (ms / 1000.0)
    # result fingerprint after warmup + steps iterations: every world size must print the same numbers (class sharding and
    # the all-reduce only change the order of a few fp32 additions)
    dsyn = tr.full_memories()[1].double()           # (collective for N > 1: the memories are sharded by class)
    gmax = tr.dynamic_syn.grad.abs().max().reshape(1)
    if world > 1:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    check = {'first_iteration': first, 'iterations': args.warmup + args.steps, 'loss_last': float(last_loss[0]),
             'dynamic_syn_sum': float(dsyn.sum()), 'dynamic_syn_sumsq': float((dsyn * dsyn).sum()),
             'hal_weight_sum': float(tr.hal.encoder.weight.detach().double().sum()),
             'grad_dynamic_absmax': float(gmax), 'lr_dynamic': LR_DYNAMIC, 'lr_hal': LR_HAL}
    del dsyn

    # ---- e2e: the step's real videos come from pinned host memory as normalised fp32 (the reference's TensorDataset,
    # get_images(...).to(device), distill_s2d_ms.py:81-87), the loss is read back.  Double-buffered: while step i computes,
    # the host->device copies of step i+1's sampled videos run on a copy stream (the sampling only depends on the numpy RNG
    # stream, not on results).  Every timed step issues exactly one full set of copies inside the timed region.
    n_own_real = C * min(BATCH_REAL, per_rank)       # upper bound of this rank's share of a draw (every class, its positions)
    e2e_steps = max(1, args.steps)
    host = torch.empty(vids.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(vids)
    bytes_video = T * 3 * HW * HW * 4
    copy_stream = torch.cuda.Stream(device=dev)
    stages = [torch.empty(n_own_real, T, 3, HW, HW, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    pending = [None, None]
    slot_box = [0]
    perm_dev = [torch.empty(n_own_real, dtype=torch.int64, device=dev) for _ in range(2)]
    perm_pin = [torch.empty(n_own_real, dtype=torch.int64).pin_memory() for _ in range(2)]
    offs_dev = [torch.empty(C + 1, dtype=torch.int32, device=dev) for _ in range(2)]
    offs_pin = [torch.empty(C + 1, dtype=torch.int32).pin_memory() for _ in range(2)]

    def make_prefetch(src_host, dst_stages):
        def prefetch(slot):
            # the sampled rows are uploaded in ascending host order with adjacent rows merged into one copy (64 of a class's
            # 72 videos are drawn, so runs are long): same bytes, ~8x fewer and larger PCIe transfers; `perm` maps sample j to
            # its row of the staging buffer, so embeddings (and the result) keep the sampled order
            real_idx = ds.sample_all_classes(BATCH_REAL)
            loc_all = ds.local_of_global[real_idx]                     # (C, n): -1 = held by another rank
            mask = loc_all >= 0
            offs = np.zeros(C + 1, dtype=np.int32)
            np.cumsum(mask.sum(1), out=offs[1:])
            loc = loc_all[mask]
            order = np.argsort(loc, kind='stable')
            srt = loc[order]
            perm = np.empty_like(order)
            perm[order] = np.arange(order.size)
            starts = np.flatnonzero(np.concatenate(([True], np.diff(srt) != 1)))
            ends = np.concatenate((starts[1:], [srt.size]))
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[slot])            # the step that last read this buffer has finished
                dst = dst_stages[slot]
                for a, b in zip(starts, ends):
                    dst[a:b].copy_(src_host[int(srt[a]):int(srt[a]) + int(b - a)], non_blocking=True)
                perm_pin[slot][:perm.size].copy_(torch.from_numpy(perm))
                perm_dev[slot].copy_(perm_pin[slot], non_blocking=True)
                offs_pin[slot].copy_(torch.from_numpy(offs))
                offs_dev[slot].copy_(offs_pin[slot], non_blocking=True)
                ready[slot].record(copy_stream)
            pending[slot] = (real_idx, int(perm.size))
        return prefetch

    def make_step(prefetch, dst_stages):
        def step_streaming():
            seed_box[0] += 1
            slot = slot_box[0]
            slot_box[0] ^= 1
            torch.cuda.current_stream().wait_event(ready[slot])
            real_idx, n_loc = pending[slot]
            loss = tr.step(net_seed=seed_box[0], real_idx=real_idx, real_batch=dst_stages[slot], real_batch_index=perm_dev[slot][:n_loc],
                           real_batch_offsets=offs_dev[slot] if world > 1 else None)
            free[slot].record()
            prefetch(slot ^ 1)                                # next step's inputs: H2D overlaps this step's kernels
            return loss.item()                                # D2H read of the step's result
        return step_streaming

    def streaming_leg(src_host, dst_stages, nbytes, note):
        prefetch = make_prefetch(src_host, dst_stages)
        step = make_step(prefetch, dst_stages)
        for ev in free:
            ev.record()
        prefetch(slot_box[0])
        step()
        ms_s = timed(step, e2e_steps)
        torch.cuda.synchronize()
        return {'value': e2e_steps / (ms_s / 1000.0), 'unit': 'it/s', 'h2d_bytes_per_step': int(nbytes), 'd2h_bytes_per_step': 4,
                'steps': e2e_steps, 'ms_per_step': ms_s / e2e_steps, 'note': note}

    e2e_fp32 = streaming_leg(host, stages, C * BATCH_REAL * bytes_video + C * BATCH_REAL * 8,
                             f'per step: {C * BATCH_REAL} sampled real videos (normalised fp32, the reference\'s preloaded TensorDataset) '
                             f'copied from pinned host memory, double-buffered on a copy stream, + loss.item()')
    del stages

    # ---- e2e (uint8 host): the host keeps the decoded uint8 frames; the normalisation (u/255 - mean)/std is fused into the
    # packer (bit-identical operands): a quarter of the PCIe bytes of the fp32 host tensors
    e2e_u8 = None
    if tr.embedder.tc is not None:
        tr.embedder.tc.set_normalization(MEAN, STD)
        stages8 = [torch.empty(n_own_real, T, 3, HW, HW, dtype=torch.uint8, device=dev) for _ in range(2)]
        e2e_u8 = streaming_leg(frames_host, stages8, C * BATCH_REAL * bytes_video // 4 + C * BATCH_REAL * 8,
                               f'host keeps the decoded uint8 frames; per step {C * BATCH_REAL} sampled videos over PCIe, normalisation fused into the packer')
        del stages8

    # ---- e2e (resident): dataset uploaded once, per-step host input = the sampled index table
    def step_resident_e2e():
        return step_resident().item()
    ms_res = timed(step_resident_e2e, e2e_steps)
    e2e_res = e2e_steps / (ms_res / 1000.0)

    # ---- per-phase device timeline of the resident step (CUDA events between the phases; max over ranks per phase)
    timeline = None
    if args.timeline:
        tr.timeline = []
        n_tl = 5
        host_t0 = time.perf_counter()
        for _ in range(n_tl):
            step_resident()
        host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / n_tl
        torch.cuda.synchronize()
        marks, tr.timeline = tr.timeline, None
        per = len(marks) // n_tl
        names = [marks[i][0] for i in range(1, per)]
        acc = torch.zeros(per - 1, device=dev)
        for k in range(n_tl):
            for i in range(1, per):
                acc[i - 1] += marks[k * per + i - 1][1].elapsed_time(marks[k * per + i][1])
        acc /= n_tl
        mx, mn = acc.clone(), acc.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        timeline = {'phases_ms_max_over_ranks': {n: float(v) for n, v in zip(names, mx)},
                    'phases_ms_min_over_ranks': {n: float(v) for n, v in zip(names, mn)},
                    'sum_ms': float(mx.sum()), 'host_enqueue_ms_per_step_rank0': host_enqueue_ms, 'steps': n_tl}

    # ---- memory-bound kernels inside the real step: CUPTI durations (torch.profiler) vs algorithmic bytes
    mem_kernels = None
    if rank == 0 and world == 1:
        try:
            mem_kernels = memory_kernel_rooflines(step_resident, tr)
        except Exception as e:                        # profiling aid only; never fail the bench line
            mem_kernels = {'error': repr(e)[:200]}

    # ---- throughput mode beside the parity mode: the single-pass bf16 pipeline on the same workload
    throughput = None
    other_modes = None
    if args.precision.startswith('f16x3') and not args.no_throughput_mode:
        def other(precision, note):
            ds.x0 = None
            torch.cuda.empty_cache()
            ds2 = copy.copy(ds)
            tr2 = make_trainer(precision, ds2)
            ds2.prepack(tr2.embedder.tc, extra_slots=len(tr2.owned) * tr2.vpc)
            for _ in range(max(2, args.warmup // 2)):
                step_resident(tr2)
            ms_t = timed(lambda: step_resident(tr2), e2e_steps)
            del tr2, ds2
            torch.cuda.empty_cache()
            return {'value': e2e_steps / (ms_t / 1000.0), 'unit': 'it/s', 'ms_per_step': ms_t / e2e_steps, 'steps': e2e_steps,
                    'dtype': precision, 'note': note}
        throughput = other('bf16', "precision='bf16': single-pass bf16 operands and bf16 activations between the layers (embeddings ~2e-3, "
                                   'unconditioned synthetic gradient ~1e-1 of fp32: NOT the parity mode)')
        if args.precision == 'f16x3r2':
            other_modes = {'f16x3': other('f16x3', "precision='f16x3': three products per MAC for the frozen real videos too (per-video real "
                                                   'embeddings 7e-5 instead of 2e-4; class means, loss and gradients as the default mode)')}

    # ---- the reference's own modules on this GPU (cuDNN), BASELINE.json configs[1] "vs reference torch-CUDA"
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        try:
            ref_cuda = reference_torch_cuda(dev, host)
        except Exception as e:
            ref_cuda = {'error': repr(e)[:300]}
    del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    l1_avg_ms = layer_ms[1] / max(1, layer_launches[1])
    l1_flops_per_launch = F_L1 * layer_videos[1] / max(1, layer_launches[1])
    achieved = l1_flops_per_launch / (l1_avg_ms * 1e-3) / 1e12 if l1_avg_ms > 0 else 0.0
    peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0)))
    split = args.precision.startswith('f16x3')
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', ('r02c_traffic.json' if real_products == 2 else 'r02_traffic.json') if split else 'r01_traffic.json')
    if os.path.exists(tpath) and (T, HW) == (16, 112):      # DRAM bytes per launch from the committed ncu --set full capture (U shape)
        tj = json.load(open(tpath))['dram_bytes_per_video']['conv1']
        traffic = (tj['read'] + tj['write']) * layer_videos[1] / max(1, layer_launches[1])
    mma_per_mac = real_products if split else 1
    roofline = {'bound': 'tensor', 'kernel': f'ws_gemm_kernel<{"EPI_L1S" if split else "EPI_L1"}> (conv 1, 64->128)', 'achieved': achieved, 'peak': peak,
                'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_kind': f'bf16_tflops_sustained of {peak_kind}',
                'avg_launch_ms': l1_avg_ms, 'flops_per_launch': l1_flops_per_launch,
                'note': (f'algorithmic FLOPs (2*M*N*K of the convolution); this launch issues {mma_per_mac} tensor-core MACs per algorithmic MAC '
                         + ('(xh*wh + xl*wh + xh*wl)' if mma_per_mac == 3 else '(xh*wh + xh*wl: frozen real videos, exact weights)')
                         + f', so the tensor pipe runs at {mma_per_mac}x this rate' if split else 'algorithmic FLOPs (2*M*N*K of the convolution)'),
                'mma_per_mac': mma_per_mac,
                'issued_tflops': achieved * mma_per_mac * (50.0 / 49.0 if split else 1.0), 'issued_frac': achieved * mma_per_mac / peak,
                'per_layer_ms_per_step': {f'conv{k}': layer_ms[k] / args.steps for k in layer_ms},
                'per_layer_tflops': {f'conv{k}': (f * layer_videos[k] / (layer_ms[k] * 1e-3) / 1e12 if layer_ms[k] > 0 else 0.0)
                                     for k, f in ((0, F_L0), (1, F_L1), (2, F_L2))}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        refdir = reference_dir()
        if refdir is not None:
            ref_utils, _ = import_reference(refdir)
            probe = ReferenceLoop(ref_utils, 'cpu', 2)
            _, dt2, _ = reference_cpu_rate(probe, threads)
            n_cls = int(max(2, min(C // 2, 15.0 / max(dt2 / 2, 1e-3))))                       # ~15 s of CPU work
            np.random.seed(0)
            its, dt, desc = reference_cpu_rate(ReferenceLoop(ref_utils, 'cpu', n_cls), threads)
            cpu = {'value': its, 'unit': 'it/s', 'cores': threads, 'kind': 'reference', 'sample': desc, 'sample_seconds': dt}
        else:
            its, dt, desc = cpu_oracle_rate(sample_classes(15.0, threads), BATCH_REAL, threads)
            cpu = {'value': its, 'unit': 'it/s', 'cores': threads, 'kind': 'port', 'sample': desc, 'sample_seconds': dt}
    dtype = {'f16x3': 'f16x3', 'f16x3r2': 'f16x3 (synthetic) / f16x2 (frozen real)', 'bf16': 'bf16'}.get(args.precision, 'f32')
    out = {
        'metric': 'DM+S2D distill iters/sec', 'value': value, 'unit': 'it/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': dtype, 'data': 'synthetic',
        'config': bench_config(world),
        'impl_detail': {'precision': args.precision,
                        'real_embed': {'f16x3': 'tcgen05, fp16 hi/lo operand pairs (3 MMAs per MAC), fp32 accumulate in TMEM',
                                       'f16x3r2': 'tcgen05, fp16 activations x fp16 hi/lo weight pairs (2 MMAs per MAC: exact weights, one rounding '
                                                  'per activation that averages out of the class mean), fp32 accumulate in TMEM; synthetic videos: '
                                                  'hi/lo pairs on both operands (3 MMAs per MAC)',
                                       'bf16': 'tcgen05 bf16 operands / fp32 accumulate'}.get(args.precision, 'fp32 CUDA cores'),
                        'syn_branch': args.syn_mode,
                        'real_set': 'resident fp32' + ('' if args.no_prepack else ' + pre-packed conv-0 operand (one-time, outside the timed region)'),
                        'prepack': prepack, 'videos_per_sec': value * C * (BATCH_REAL + VPC)},
        'clocks': clocks, 'gpu_launches': int(launches), 'check': check,
        # headline end-to-end number: the reference's own host format (normalised fp32 TensorDataset)
        'e2e': e2e_fp32,
        'e2e_uint8_host': e2e_u8,
        'e2e_resident': {'value': e2e_res, 'unit': 'it/s', 'h2d_bytes_per_step': int(C * BATCH_REAL * 8),
                         'd2h_bytes_per_step': 4, 'steps': e2e_steps,
                         'note': 'real set uploaded once; per step the host sends the sampled index table and reads the loss'},
        'throughput_mode': throughput, 'other_modes': other_modes, 'reference_torch_cuda': ref_cuda, 'timeline': timeline,
        'roofline': roofline, 'memory_kernels': mem_kernels, 'cpu_baseline': cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ MTT + S2D (configs[3])
def mtt_flops_per_iteration():
    """Algorithmic FLOPs of one MTT+S2D iteration, BASELINE.md section 3: "<= 45.3 TFLOP" for syn_steps = 10 x 50 videos of the
    U shape (per step and video: fprop + dgrad + wgrad of the inner gradient and their second-order counterparts in
    grand_loss.backward(), ~8.2 conv-equivalents of 11.002 GFLOP), scaled linearly to other shapes."""
    return 45.3e12 * (F_EMBED / 11.002e9) * (SYN_STEPS / 10.0) * (C * VPC / 50.0)


def mtt_config(n_gpus):
    return {'workload': WORKLOAD_DESC,
            'parallelism': f'every unrolled step batch sharded i%{n_gpus}, flat-parameter gradient all-reduced (NCCL)' if n_gpus > 1 else 'single GPU',
            'syn_steps': SYN_STEPS, 'batch_syn': C * VPC,
            'l2_policy': 'each unrolled step re-packs activations and weights (> L2); 14.6 MB flat parameters per expert snapshot'}


class ReferenceMTTLoop:
    """distill_s2d_ms.py:89-108 state and the VERBATIM MTT iteration body (:196-300) through the reference's own modules
    (utils.get_network, reparam_module.ReparamModule, utils.Conv3DNet, torch.optim.SGD).  `batch_syn` / `syn_steps` smaller
    than the workload's give a bounded sample (cost is linear in steps x videos)."""

    def __init__(self, ref_utils, ref_reparam, device, syn_steps, batch_syn):
        self.u, self.rp, self.device, self.syn_steps, self.batch_syn = ref_utils, ref_reparam, device, syn_steps, batch_syn
        torch.manual_seed(0)
        self.static_syn = torch.randn(size=(C * SPC, 3, HW, HW), dtype=torch.float).detach().to(device).requires_grad_(False)
        self.dynamic_syn = torch.randn(size=(C, DPC, T, 1, HW, HW), dtype=torch.float).detach().to(device).requires_grad_(True)
        self.hals = torch.nn.ModuleList([ref_utils.Conv3DNet()]).to(device)
        self.syn_lr = torch.tensor(0.01).detach().to(device).requires_grad_(True)
        self.optimizer_dynamic = torch.optim.SGD([self.dynamic_syn], lr=LR_DYNAMIC, momentum=0.95)
        self.optimizer_hals = torch.optim.SGD(self.hals.parameters(), lr=LR_HAL, momentum=0.95)
        self.optimizer_lr = torch.optim.SGD([self.syn_lr], lr=1e-5, momentum=0.9)
        self.criterion = torch.nn.CrossEntropyLoss().to(device)
        base = ref_utils.get_network('ConvNet3D', 3, C, (HW, HW), frames=T, dist=False)
        self.start = [p.detach().clone() for p in base.parameters()]
        self.target = [p.detach().clone() + 0.01 * torch.randn_like(p) for p in base.parameters()]

    def iteration(self):
        dev, num_classes, vpc, spc = self.device, C, VPC, SPC
        student_net = self.u.get_network('ConvNet3D', 3, num_classes, (HW, HW), frames=T, dist=False).to(dev)
        student_net = self.rp.ReparamModule(student_net)
        student_net.train()
        num_params = sum([np.prod(p.size()) for p in (student_net.parameters())])
        target_params = torch.cat([p.data.to(dev).reshape(-1) for p in self.target], 0)
        student_params = [torch.cat([p.data.to(dev).reshape(-1) for p in self.start], 0).requires_grad_(True)]
        starting_params = torch.cat([p.data.to(dev).reshape(-1) for p in self.start], 0)
        indices_chunks = []
        for step in range(self.syn_steps):
            if not indices_chunks:
                indices = torch.randperm(num_classes * vpc, device=dev)
                indices_chunks = list(torch.split(indices, self.batch_syn))
            these_indices = indices_chunks.pop()
            label = these_indices // vpc
            idx = these_indices % vpc
            dynamic_idx = 2 * idx + torch.randint(2, (these_indices.shape[0],), device=dev)
            static_idx = spc * label + 2 * idx + torch.randint(2, (these_indices.shape[0],), device=dev)
            static = self.static_syn[static_idx, :, :, :]
            dynamic = self.dynamic_syn[label, dynamic_idx, :, :, :, :]
            x = self.hals[0](static, dynamic)
            this_y = label.long()
            x = student_net(x, flat_param=student_params[-1])
            loss = self.criterion(x, this_y)
            grad = torch.autograd.grad(loss, student_params[-1], create_graph=True)[0]
            student_params.append(student_params[-1] - self.syn_lr * grad)
        param_loss = torch.nn.functional.mse_loss(student_params[-1], target_params, reduction='sum')
        param_dist = torch.nn.functional.mse_loss(starting_params, target_params, reduction='sum')
        param_loss = param_loss / num_params
        param_dist = param_dist / num_params
        grand_loss = param_loss / param_dist
        self.optimizer_dynamic.zero_grad()
        self.optimizer_hals.zero_grad()
        self.optimizer_lr.zero_grad()
        grand_loss.backward()
        self.optimizer_dynamic.step()
        self.optimizer_hals.step()
        self.optimizer_lr.step()
        self.syn_lr.data = self.syn_lr.data.clip(min=0.001)
        return grand_loss.item()


def import_reference_reparam(path):
    saved = list(sys.path)
    hidden = {n: sys.modules.pop(n) for n in ('reparam_module',) if n in sys.modules}
    sys.path[:] = [path] + [q for q in saved if os.path.abspath(q or '.') != ROOT]
    try:
        import reparam_module as ref_reparam          # noqa
    finally:
        sys.modules.pop('reparam_module', None)
        sys.modules.update(hidden)
        sys.path[:] = saved
    return ref_reparam


def reference_mtt_cpu_rate(threads, syn_steps=2, batch_syn=None):
    """One reference MTT iteration on the host cores over a bounded sample: the workload's true step batch (50 videos: the CPU cost
    per video depends on the batch) with syn_steps of its 10 unrolled steps; returns (it/s scaled linearly in the unroll depth,
    seconds, description) or None when the reference modules are absent."""
    refdir = reference_dir()
    if refdir is None:
        return None
    if batch_syn is None:
        batch_syn = C * VPC
    ref_utils, _ = import_reference(refdir)
    torch.set_num_threads(threads)
    loop = ReferenceMTTLoop(ref_utils, import_reference_reparam(refdir), 'cpu', syn_steps, batch_syn)
    t0 = time.perf_counter()
    loss = loop.iteration()
    dt = time.perf_counter() - t0
    assert np.isfinite(loss)
    scale = (SYN_STEPS * C * VPC) / (syn_steps * batch_syn)
    desc = (f'{syn_steps} of {SYN_STEPS} unrolled steps x {batch_syn} of {C * VPC} videos {T}x3x{HW}x{HW} through the reference\'s own modules '
            f'(ReparamModule student, loop body of distill_s2d_ms.py:196-300, incl. grand_loss.backward()), time scaled by {scale:.1f}')
    return 1.0 / (dt * scale), dt, desc


def run_mtt_reference(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1
    vals, secs, desc = [], [], ''
    for i in range(max(1, args.warmup + args.steps)):
        r = reference_mtt_cpu_rate(threads)
        if r is None:
            print(json.dumps({'impl': 'reference', 'unavailable': 'reference modules not found (baseline/_ref, $VD_REFERENCE, /root/reference)'}))
            return
        if i >= args.warmup:
            vals.append(r[0]); secs.append(r[1]); desc = r[2]
    v = float(np.mean(vals))
    print(json.dumps({'impl': 'reference', 'metric': 'MTT+S2D distill iters/sec', 'value': v, 'unit': 'it/s', 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * float(np.mean(secs)), 'ms_per_full_iteration': 1000.0 / v,
                      'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': mtt_config(args.gpus),
                      'cpu_baseline': {'value': v, 'unit': 'it/s', 'cores': threads, 'kind': 'reference', 'sample': desc, 'sample_seconds': float(np.mean(secs))},
                      'e2e': {'value': v, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def run_mtt_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', ''):
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from video_distillation_b200 import _lib
    from video_distillation_b200.distill import MTTS2DTrainer
    from video_distillation_b200.networks import ConvNet3D
    # default: the parity-grade tensor-core trio (every fprop / dgrad / wgrad on hi / lo operand pairs: <= 1e-3 of the reference on
    # every gradient, tests/test_dm_gpu.py::test_mtt_s2d_golden[bf16x3]); --precision bf16 = split fprop + single-pass dgrad / wgrad
    precision = {'fp32': 'fp32', 'bf16': 'bf16'}.get(args.precision, 'bf16x3')
    torch.manual_seed(0)
    tr = MTTS2DTrainer(num_classes=C, im_size=(HW, HW), frames=T, vpc=VPC, spc=SPC, dpc=DPC, syn_steps=SYN_STEPS, lr_dynamic=LR_DYNAMIC,
                       lr_hal=LR_HAL, device=dev, precision=precision)
    base = ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW))
    # synthetic expert segment: (start, target) parameter snapshots in pinned host memory, like a loaded replay buffer
    start = [p.detach().clone().pin_memory() for p in base.parameters()]
    target = [(p.detach() + 0.01 * torch.randn_like(p)).pin_memory() for p in base.parameters()]
    start_dev, target_dev = [p.to(dev) for p in start], [p.to(dev) for p in target]
    h2d = 2 * sum(p.numel() * 4 for p in start)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()
    it_box = [0]
    last = [None]

    def step_resident():
        it_box[0] += 1
        last[0] = tr.step(start_dev, target_dev, net_seed=it_box[0])
        return last[0]

    def step_e2e():                        # expert snapshots come from (pinned) host memory every iteration, the grand loss is read back
        it_box[0] += 1
        return tr.step([p.to(dev, non_blocking=True) for p in start], [p.to(dev, non_blocking=True) for p in target], net_seed=it_box[0]).item()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    _lib.launch_count_reset()
    t_wall0 = time.time()
    ms = timed(step_resident, args.steps)
    sampler.mark(t_wall0, time.time())
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps / (ms / 1000.0)
    check = {'iterations': args.warmup + args.steps, 'grand_loss_last': float(last[0]), 'syn_lr': float(tr.syn_lr.detach()),
             'dynamic_syn_sumsq': float((tr.dynamic_syn.detach().double() ** 2).sum())}
    ms_e = timed(step_e2e, args.steps)
    throughput_mode = None
    if precision == 'bf16x3' and not args.no_throughput_mode:
        torch.manual_seed(0)
        tr_main, tr = tr, MTTS2DTrainer(num_classes=C, im_size=(HW, HW), frames=T, vpc=VPC, spc=SPC, dpc=DPC, syn_steps=SYN_STEPS,
                                        lr_dynamic=LR_DYNAMIC, lr_hal=LR_HAL, device=dev, precision='bf16')
        for _ in range(2):
            step_resident()
        ms_t = timed(step_resident, args.steps)
        throughput_mode = {'value': args.steps / (ms_t / 1000.0), 'unit': 'it/s', 'ms_per_step': ms_t / args.steps, 'steps': args.steps,
                           'dtype': 'bf16x3 fprop / bf16 dgrad+wgrad',
                           'note': "precision='bf16': split fprop, single-pass bf16 dgrad / wgrad (hallucinator / lr gradients 3e-3 / 7e-3 / "
                                   "5e-4, dynamic memory 9e-3 of the reference: NOT the parity mode)"}
        tr = tr_main
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = measured_peaks()
    peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0)))
    flops = mtt_flops_per_iteration()
    achieved = flops / (ms / args.steps * 1e-3) / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = reference_mtt_cpu_rate(os.cpu_count() or 1)
        if r is not None:
            cpu = {'value': r[0], 'unit': 'it/s', 'cores': os.cpu_count() or 1, 'kind': 'reference', 'sample': r[2], 'sample_seconds': r[1]}
    print(json.dumps({
        'metric': 'MTT+S2D distill iters/sec', 'value': value, 'unit': 'it/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': {'bf16': 'bf16x3 fprop / bf16 dgrad+wgrad', 'bf16x3': 'bf16x3 (hi/lo operand pairs, three products per MAC, fp32 accumulate)',
                  'fp32': 'f32'}[precision], 'data': 'synthetic', 'config': mtt_config(world),
        'impl_detail': {'precision': precision, 'student': 'ReparamModule over flat parameters; conv trio (fprop / dgrad / wgrad, closed under '
                        'differentiation) on tcgen05 GEMMs' if precision != 'fp32' else 'exact fp32 CUDA-core conv trio',
                        'mma_per_mac': {'bf16x3': 3, 'bf16': '3 fprop / 1 dgrad, wgrad', 'fp32': 0}[precision]},
        'throughput_mode': throughput_mode,
        'clocks': clocks, 'gpu_launches': int(launches), 'check': check,
        'e2e': {'value': args.steps / (ms_e / 1000.0), 'unit': 'it/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4, 'steps': args.steps,
                'note': 'expert start / target snapshots copied from pinned host memory every iteration + grand_loss.item()'},
        'roofline': {'bound': 'tensor', 'kernel': 'whole iteration (the conv-trio GEMMs ws_gemm_kernel<...> are ~all of it)', 'achieved': achieved,
                     'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None, 'peak_kind': f'bf16_tflops_sustained of {peak_kind}',
                     'flops_per_iteration': flops,
                     'note': 'algorithmic FLOPs of the unroll (BASELINE.md section 3, <= 45.3 TFLOP) / device time of the iteration; the trio still '
                             'materialises im2col (wgrad) and column buffers (dgrad of conv 2) in HBM'},
        'cpu_baseline': cpu}))
    if world > 1:
        dist.destroy_process_group()


WORKLOAD_DESC = ''
WORKLOADS = {   # name -> (C, T, HW, vpc, spc, dpc, batch_real, per_class, FLOP per video L0/L1/L2, BASELINE.json config)
    'U-ipc1': (50, 16, 112, 1, 2, 2, 64, 72, (2.832e9, 7.553e9, 0.617e9), 'configs[1]'),
    'U-ipc5': (50, 16, 112, 5, 10, 10, 64, 72, (2.832e9, 7.553e9, 0.617e9), 'configs[2]'),
    'K-ipc5': (400, 8, 64, 5, 10, 10, 64, 64, (0.462e9, 1.233e9, 0.077e9), 'configs[4]'),
    # MTT + S2D (BASELINE.json configs[3]): batch_real / per_class unused; syn_steps = 10 unrolled student steps of C*vpc videos
    'MTT-U-ipc1': (50, 16, 112, 1, 2, 2, 0, 0, (2.832e9, 7.553e9, 0.617e9), 'configs[3]'),
}
SYN_STEPS = 10
MTT = False


def set_workload(name):
    global C, T, HW, VPC, SPC, DPC, BATCH_REAL, PER_CLASS, F_L0, F_L1, F_L2, F_EMBED, WORKLOAD_DESC, MTT
    C, T, HW, VPC, SPC, DPC, BATCH_REAL, PER_CLASS, (F_L0, F_L1, F_L2), cfg = WORKLOADS[name]
    F_EMBED = F_L0 + F_L1 + F_L2
    MTT = name.startswith('MTT')
    if MTT:
        WORKLOAD_DESC = (f'MTT+S2D miniUCF101-shape: {C} classes, videos {T}x3x{HW}x{HW}, vpc={VPC} spc={SPC} dpc={DPC}, syn_steps={SYN_STEPS} '
                         f'unrolled ReparamModule student steps of batch_syn={C * VPC} videos against a synthetic expert segment (BASELINE.json {cfg})')
        return
    WORKLOAD_DESC = (f'DM+S2D {"miniUCF101" if HW == 112 else "Kinetics-400"}-shape: {C} classes, videos {T}x3x{HW}x{HW}, vpc={VPC} spc={SPC} '
                     f'dpc={DPC}, batch_real={BATCH_REAL} (BASELINE.json {cfg}); fresh frozen ConvNet3D per step')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='f16x3r2', choices=['f16x3r2', 'f16x3', 'bf16', 'fp32'],
                    help='f16x3r2 (default): fused tcgen05 pipeline on fp16 hi/lo operand pairs, the parity mode — three products per MAC '
                         'for the synthetic videos, two (exact weights, fp16 activations) for the frozen real videos; f16x3: three products '
                         'everywhere; bf16: single-pass throughput mode; fp32: exact CUDA-core kernels')
    ap.add_argument('--timeline', action='store_true', help='add a per-phase device timeline of the step (CUDA events, max / min over ranks)')
    ap.add_argument('--no-throughput-mode', action='store_true', help='skip the single-pass bf16 line reported beside the default mode')
    ap.add_argument('--no-reference-cuda', action='store_true', help="skip the reference's own modules on this GPU (cuDNN)")
    ap.add_argument('--max-batch', type=int, default=640)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--syn-mode', default='fused', choices=['fused', 'split', 'fp32'],
                    help='synthetic branch: fused bf16 tensor-core pipeline (throughput), split-bf16 trio with fp32 '
                         'activations (gradients within 1e-2 of fp32), or the exact fp32 CUDA-core kernels')
    ap.add_argument('--no-prepack', action='store_true', help='pack the sampled real videos every step instead of once')
    ap.add_argument('--workload', default='U-ipc1', choices=sorted(WORKLOADS),
                    help='U-ipc1 = the metric workload (BASELINE.json configs[1], default); U-ipc5 / K-ipc5 = configs[2] / configs[4] '
                         '(parity-test shapes, timed on request)')
    args = ap.parse_args()
    set_workload(args.workload)
    if MTT:
        run_mtt_reference(args) if args.impl == 'reference' else run_mtt_ours(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
