"""Where does an MMA issuer thread's time go?  Per-layer cycle counters (vd_tc_set_profile_buffer)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _probe_lib import use_probe_library  # noqa: E402
use_probe_library()               # probe entry points live in scripts/libvd_b200_probe.so, not in the product library
from video_distillation_b200 import _lib  # noqa: E402
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B, T, HW = int(os.environ.get('B', 592)), int(os.environ.get('T', 16)), int(os.environ.get('HW', 112))
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
SPLIT = len(sys.argv) > 1 and sys.argv[1] in ('x3', 'x2')
PRODUCTS = 2 if (len(sys.argv) > 1 and sys.argv[1] == 'x2') else 3
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=B, split=SPLIT, real_products=PRODUCTS)
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(64, T, 3, HW, HW, device='cuda')
idx = torch.arange(B, device='cuda') % 64
x0 = tc.pack_dataset(video)
tc.embed_resident(x0, idx)
buf = torch.zeros(148 * 8, dtype=torch.int64, device='cuda')
lib = _lib.lib()
a1, a2 = tc._buffers(B)
out = torch.empty(B, tc.embed_dim, device='cuda')
lib.vd_tc_set_profile_buffer(_lib.ptr(buf))
LAYERS = [int(v) for v in os.environ.get('LAYERS', '0,1,2').split(',')]
for layer, (src, w, b, dst, ii) in enumerate([(x0, tc.w0, tc.b0, a1, idx), (a1, tc.w1, tc.b1, a2, None), (a2, tc.w2, tc.b2, out, None)]):
    if layer not in LAYERS:
        continue
    buf.zero_()
    tc._conv_layer(layer, src, w, b, dst, B, None, ii, False, 0, PRODUCTS)
    torch.cuda.synchronize()
    v = buf.cpu().view(148, 8).double().mean(0)
    tot = v[0].item()
    # counters of MMA issuer 0 (issuer 1 mirrors it): o[1..3] = time blocked on acc_empty / pix_full / w_full for ITS groups,
    # o[4] = fence + MMA issue + baton arrive + weight-slot commit, o[5] = time waiting for the baton of the other issuer
    print(f'conv{layer}: total {tot:10.0f} cyc | wait acc_empty {100 * v[1] / tot:5.1f}% | wait pix_full {100 * v[2] / tot:5.1f}% | '
          f'wait w_full {100 * v[3] / tot:5.1f}% | issue + baton {100 * v[4] / tot:5.1f}% | '
          f'other (generator, commits of the peer\'s groups) {100 * (tot - v[1:5].sum().item()) / tot:5.1f}% || epilogue warp 0: '
          f'wait acc_full {100 * v[5] / tot:5.1f}% drain {100 * v[6] / tot:5.1f}% store {100 * v[7] / tot:5.1f}%')
lib.vd_tc_set_profile_buffer(None)
