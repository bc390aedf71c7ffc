"""Portable synthetic tensors for golden vectors (TEST INFRASTRUCTURE ONLY).

Golden inputs are NOT stored: they are regenerated bit-exactly from integer hashes (numpy
uint64 arithmetic), so fixtures only hold the reference outputs.  The value grid is coarse
(multiples of 2^-8 in [-1, 1)) so every input is exactly representable in bf16 as well.
"""
import math

import numpy as np
import torch


def hash_uniform(shape, seed, scale=1.0):
    """float32 tensor of ``shape`` with values k/256, k in [-256, 256), from a 64-bit mix hash."""
    n = int(np.prod(shape))
    with np.errstate(over='ignore'):
        return _hash_uniform(shape, n, seed, scale)


def _hash_uniform(shape, n, seed, scale):
    i = np.arange(n, dtype=np.uint64)
    x = (i + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(30)
    x = (x * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(27)
    x = (x * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(31)
    k = (x >> np.uint64(55)).astype(np.int64) - 256          # 9 bits -> [-256, 256)
    v = (k.astype(np.float32) / np.float32(256.0)) * np.float32(scale)
    return torch.from_numpy(v.reshape(shape))


def synth_convnet3d_params(seed, channel=3, num_classes=50, net_width=128, net_depth=3):
    """ConvNet3D-shaped parameters (net_norm='none') with hash values scaled like the default
    init bound 1/sqrt(fan_in) (power-of-two rounded so values stay bf16-exact)."""
    params = {}
    cin = channel
    for d in range(net_depth):
        cout = 64 if d == 0 else net_width
        fan_in = cin * 147
        s = 2.0 ** math.floor(math.log2(1.0 / math.sqrt(fan_in)))
        params[f'features.{3 * d}.weight'] = hash_uniform((cout, cin, 3, 7, 7), seed * 100 + 2 * d, s)
        params[f'features.{3 * d}.bias'] = hash_uniform((cout,), seed * 100 + 2 * d + 1, s)
        cin = cout
    s = 2.0 ** math.floor(math.log2(1.0 / math.sqrt(net_width)))
    params['logit.weight'] = hash_uniform((num_classes, net_width, 1, 1, 1), seed * 100 + 50, s)
    params['logit.bias'] = hash_uniform((num_classes,), seed * 100 + 51, s)
    return params


def synth_hallucinator(seed):
    return {'encoder.weight': hash_uniform((3, 4, 3, 3, 3), seed * 100 + 60, 0.125),
            'encoder.bias': hash_uniform((3,), seed * 100 + 61, 0.125)}


def summarize(t, stride=97):
    """Compact fingerprint of a big tensor: (sum, sum of squares, strided sample) in float64."""
    f = t.detach().reshape(-1).double()
    return np.array([f.sum().item(), (f * f).sum().item()]), f[::stride].numpy().copy()
