"""Oracle restatement of the static-dynamic composer (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils.py:1178-1197 (class Conv3DNet, the "hallucinator").
Parity pinning: oracle/make_golden.py -> tests/golden/composer_*.npz.
"""
import math

import torch
import torch.nn.functional as F


def init_hallucinator(seed=None, mode='concat'):
    """``Conv3DNet()`` default construction (utils.py:1179-1184): one nn.Conv3d(4|3 -> 3, k=3, pad=1)
    with torch's default init.  ``seed`` (optional) seeds the global generator first."""
    if seed is not None:
        torch.random.manual_seed(seed)
    cin = 4 if mode == 'concat' else 3
    w = torch.empty(3, cin, 3, 3, 3)
    b = torch.empty(3)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(cin * 27)
    torch.nn.init.uniform_(b, -bound, bound)
    return {'encoder.weight': w, 'encoder.bias': b}


def compose(static, dynamic, weight, bias, mode='concat'):
    """Conv3DNet.forward (utils.py:1186-1197).

    static  (B, 3, H, W)      still image ("static memory")
    dynamic (B, T, 1, H, W)   per-frame residual ("dynamic memory")
    returns (B, T, 3, H, W):  conv3d over cat([static broadcast over T, dynamic]) (concat mode)
    or over static + dynamic (add mode).
    """
    b, f, _, h, w = dynamic.shape
    s = static.unsqueeze(2).expand(b, 3, f, h, w)          # repeat over T, utils.py:1188
    d = dynamic.permute(0, 2, 1, 3, 4)                      # (B,1,T,H,W), utils.py:1189
    if mode == 'concat':
        x = torch.cat([s, d], dim=1)
    elif mode == 'add':
        x = s + d
    else:
        raise NotImplementedError
    y = F.conv3d(x, weight, bias, padding=1)
    return y.permute(0, 2, 1, 3, 4)
