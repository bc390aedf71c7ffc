"""Which host calls of one DM+S2D iteration synchronise with the device?  (torch sync debug mode)"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer  # noqa: E402

C, T, HW, PER = 7, 16, 112, 72
dev = torch.device('cuda', 0)
labels = [c for c in range(C) for _ in range(PER)]
vids = torch.randn(C * PER, T, 3, HW, HW, device=dev)
ds = DeviceDataset.from_device_shard(vids, labels, C, dev, 0, 1)
tr = DMS2DTrainer(ds, num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, batch_real=64, precision='bf16',
                  device=dev, init_on_device=True, max_batch=640)
ds.prepack(tr.embedder.tc, extra_slots=C)
np.random.seed(0)
for i in range(3):
    tr.step(net_seed=i)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode('warn')
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    tr.step(net_seed=5)
    for x in w:
        print('SYNC:', str(x.message)[:200], '@', x.filename.split('/')[-1], x.lineno)
torch.cuda.set_sync_debug_mode('default')
import traceback
print('done;', len(w), 'synchronising calls')
