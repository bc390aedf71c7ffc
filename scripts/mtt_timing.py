"""One MTT+S2D iteration at the miniUCF101 shape (BASELINE.json configs[3]): syn_steps unrolled student steps
of batch_syn = C*vpc videos, grand loss, backward to the synthetic memories."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.distill import MTTS2DTrainer  # noqa: E402
from video_distillation_b200.networks import ConvNet3D  # noqa: E402

world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
if world > 1:                                   # torchrun: every inner step's 50 videos are sharded over the ranks
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
C, T, HW = 50, 16, 112
syn_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
torch.manual_seed(0)
tr = MTTS2DTrainer(num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, syn_steps=syn_steps, device='cuda', precision=precision)
base = ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW))
start = [p.detach().clone() for p in base.parameters()]
target = [p.detach().clone() + 0.01 * torch.randn_like(p) for p in base.parameters()]
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = tr.step(start, target, net_seed=1)
    torch.cuda.synchronize()
    if rank == 0:
        print(f'MTT iteration {i} [{precision}]: syn_steps={syn_steps} batch_syn={C} -> {time.perf_counter() - t0:.3f} s, grand loss {loss.item():.6f}, '
              f'world {world}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB', flush=True)
if os.environ.get('VD_MTT_PROFILE') and rank == 0:      # kernel-level breakdown of one iteration (CUPTI via torch.profiler)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(start, target, net_seed=1)
        torch.cuda.synchronize()
    dur = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            d = dur.setdefault(e.name[:100], [0, 0.0])
            d[0] += 1
            d[1] += (e.time_range.end - e.time_range.start) * 1e-3
    tot = sum(v[1] for v in dur.values())
    print(f'sum of kernel time {tot:.1f} ms over {sum(v[0] for v in dur.values())} launches')
    for k, v in sorted(dur.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f'{v[1]:9.2f} ms {100 * v[1] / tot:5.1f}% {v[0]:5d}x  {k}')
if world > 1:
    dist.destroy_process_group()
