"""Oracle restatement of the reference ConvNet3D (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/networks.py:727-814 (class ConvNet3D) and the factory
branch /root/reference/utils.py:518-520,608-609 (get_network('ConvNet3D')).
Everything is expressed with explicit tensors + torch.nn.functional on CPU so
that the CUDA path can be compared layer by layer.

Parity pinning: checked against the live reference modules by
oracle/make_golden.py -> tests/golden/convnet3d_*.npz.
"""
import math

import torch
import torch.nn.functional as F

KERNEL = (3, 7, 7)   # networks.py:799
STRIDE = (1, 2, 2)
PADDING = (1, 3, 3)


def _conv_channels(channel, net_width, net_depth):
    """(cin, cout) of each feature conv: first layer is hard-wired to 64 (networks.py:799)."""
    out = []
    cin = channel
    for d in range(net_depth):
        cout = 64 if d == 0 else net_width
        out.append((cin, cout))
        cin = cout
    return out


def convnet3d_param_names(net_depth=3, net_norm='none'):
    """state_dict / ReparamModule ordering (reparam_module.py:30-41).

    With net_norm='none' each block is [conv, relu, pool] -> conv index 3*d;
    with a norm layer it is [conv, norm, relu, pool] -> conv 4*d, norm 4*d+1.
    """
    names = []
    per = 3 if net_norm == 'none' else 4
    for d in range(net_depth):
        names += [f'features.{per * d}.weight', f'features.{per * d}.bias']
        if net_norm != 'none':
            names += [f'features.{per * d + 1}.weight', f'features.{per * d + 1}.bias']
    names += ['logit.weight', 'logit.bias']
    return names


def _default_conv_init(weight, bias):
    """nn.Conv3d.reset_parameters (torch/nn/modules/conv.py): kaiming_uniform(a=sqrt 5) then
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for the bias; this is what networks.py:799,736 get."""
    torch.nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    fan_in = weight.shape[1] * weight.shape[2] * weight.shape[3] * weight.shape[4]
    bound = 1.0 / math.sqrt(fan_in)
    torch.nn.init.uniform_(bias, -bound, bound)


def init_convnet3d(seed, channel=3, num_classes=50, net_width=128, net_depth=3, net_norm='none'):
    """Parameters of get_network('ConvNet3D') under ``torch.random.manual_seed(seed)``.

    utils.py:519 seeds the global generator (from the wall clock) and then constructs the
    convs in order features.0, features.3, features.6, logit (networks.py:731,736); each
    Conv3d draws its weight then its bias.  Returns an ordered dict name -> tensor.
    """
    torch.random.manual_seed(seed)
    params = {}
    per = 3 if net_norm == 'none' else 4
    for d, (cin, cout) in enumerate(_conv_channels(channel, net_width, net_depth)):
        w = torch.empty(cout, cin, *KERNEL)
        b = torch.empty(cout)
        _default_conv_init(w, b)
        params[f'features.{per * d}.weight'] = w
        params[f'features.{per * d}.bias'] = b
        if net_norm != 'none':
            # nn.GroupNorm(C, C, affine=True): ones / zeros, no RNG (networks.py:784)
            params[f'features.{per * d + 1}.weight'] = torch.ones(cout)
            params[f'features.{per * d + 1}.bias'] = torch.zeros(cout)
    w = torch.empty(num_classes, net_width, 1, 1, 1)
    b = torch.empty(num_classes)
    _default_conv_init(w, b)
    params['logit.weight'] = w
    params['logit.bias'] = b
    return params


def flatten_params(params, names=None):
    """ReparamModule flat layout: torch.cat([p.reshape(-1)]) in registration order
    (reparam_module.py:51)."""
    names = names or list(params.keys())
    return torch.cat([params[n].reshape(-1) for n in names], 0)


def unflatten_params(flat, like):
    """Views of ``flat`` with the shapes of ``like`` (reparam_module.py:110-115)."""
    out = {}
    ofs = 0
    for n, p in like.items():
        out[n] = flat[ofs:ofs + p.numel()].view(p.shape)
        ofs += p.numel()
    assert ofs == flat.numel()
    return out


def _pool(x, net_pooling, d):
    if net_pooling == 'maxpooling':          # networks.py:766-770 (flag = 1 only for d == 0)
        k = (1, 2, 2) if d == 0 else (2, 2, 2)
        return F.max_pool3d(x, kernel_size=k, stride=k)
    if net_pooling == 'avgpooling':          # networks.py:771-772
        return F.avg_pool3d(x, kernel_size=2, stride=2)
    if net_pooling == 'none':
        return x
    raise ValueError(net_pooling)


def convnet3d_features(params, x, net_depth=3, net_norm='none', net_pooling='maxpooling',
                       return_intermediates=False):
    """``self.features`` on an NCDHW tensor (networks.py:792-814)."""
    per = 3 if net_norm == 'none' else 4
    inter = []
    for d in range(net_depth):
        w = params[f'features.{per * d}.weight']
        b = params[f'features.{per * d}.bias']
        x = F.conv3d(x, w, b, stride=STRIDE, padding=PADDING)
        if return_intermediates:
            inter.append(('conv', d, x))
        if net_norm == 'instancenorm':       # GroupNorm(C, C, affine) networks.py:784
            g = params[f'features.{per * d + 1}.weight']
            be = params[f'features.{per * d + 1}.bias']
            x = F.group_norm(x, x.shape[1], g, be, eps=1e-5)
        elif net_norm != 'none':
            raise ValueError(net_norm)
        x = F.relu(x)                         # networks.py:757
        x = _pool(x, net_pooling, d)
        if return_intermediates:
            inter.append(('pool', d, x))
    return (x, inter) if return_intermediates else x


def convnet3d_embed(params, video, **kw):
    """ConvNet3D.embed (networks.py:747-751): video is (B, T, C, H, W)."""
    x = video.permute(0, 2, 1, 3, 4)
    out = convnet3d_features(params, x, **kw)
    return out.reshape(out.size(0), -1)


def convnet3d_forward(params, video, im_size, dropout_mask=None, **kw):
    """ConvNet3D.forward (networks.py:738-745).

    ``dropout_mask``: None = eval mode (identity); otherwise a {0,1} tensor shaped like the
    avg-pooled features, applied as ``x * mask / (1 - p)`` with p = 0.5 (nn.Dropout(0.5),
    networks.py:735) so that a test can reuse the very mask the CUDA path drew.
    """
    x = video.permute(0, 2, 1, 3, 4)
    out = convnet3d_features(params, x, **kw)
    k = (2, 2, 2) if im_size[0] > 64 else (2, 1, 1)          # networks.py:733
    out = F.avg_pool3d(out, kernel_size=k, stride=(1, 1, 1))
    if dropout_mask is not None:
        out = out * dropout_mask / 0.5
    out = F.conv3d(out, params['logit.weight'], params['logit.bias'])
    logits = out.squeeze(3).squeeze(3)
    return torch.max(logits, 2)[0]


def embed_dim(frames, im_size, net_width=128, net_depth=3, net_pooling='maxpooling'):
    t, h, w = frames, im_size[0], im_size[1]
    for d in range(net_depth):
        h, w = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        if net_pooling == 'maxpooling':
            if d != 0:
                t //= 2
            h //= 2
            w //= 2
        elif net_pooling == 'avgpooling':
            t //= 2
            h //= 2
            w //= 2
    return net_width * t * h * w


def routing_codes(params, video, dtype=torch.float32):
    """ReLU masks / MaxPool argmax of the oracle forward (networks.py:757,766-770 with get_network's none/maxpooling setting) in
    the CUDA library's routing-code format: per pooled output, ``arg | active << 3`` with arg the window position in (t,h,w)
    scan order.  Lets a test impose the ORACLE's routing on the CUDA backward (routing-conditioned parity).
    Returns [code0 (B,64,T,H1,H1), code1, code2] (uint8)."""
    h = video.permute(0, 2, 1, 3, 4).to(dtype)
    codes = []
    for d in range(3):
        y = F.conv3d(h, params[f'features.{3 * d}.weight'].to(dtype), params[f'features.{3 * d}.bias'].to(dtype), STRIDE, PADDING)
        k = (1, 2, 2) if d == 0 else (2, 2, 2)
        h, idx = F.max_pool3d(F.relu(y), k, k, return_indices=True)
        To, Ho, Wo = y.shape[2:]
        it, ih, iw = idx // (Ho * Wo), (idx // Wo) % Ho, idx % Wo
        pos = (it % k[0]) * (k[1] * k[2]) + (ih % k[1]) * k[2] + (iw % k[2])
        codes.append((pos | ((h > 0).long() << 3)).to(torch.uint8))
    return codes
