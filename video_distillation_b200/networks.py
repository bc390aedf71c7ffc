"""ConvNet3D with the reference's constructor, attributes and state_dict layout
(/root/reference/networks.py:727-814) whose tensor work runs in libvd_b200 (sm_100a CUDA).

Parameter names / order are identical to the reference (features.{0,3,6}.weight|bias, logit.*;
with a norm layer features.{0,4,8} + features.{1,5,9}), so ReparamModule flattening, expert
trajectories and checkpoints are interchangeable.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

KERNEL, STRIDE, PADDING = (3, 7, 7), (1, 2, 2), (1, 3, 3)


class Conv3d(nn.Conv3d):
    """nn.Conv3d (same init, same parameters) computing through vd_conv3d_*_f32."""

    def forward(self, x):
        return ops.conv3d(x, self.weight, self.bias, self.stride, self.padding)


class ReLU(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return ops.relu_maxpool3d(x, (1, 1, 1))


class MaxPool3d(nn.Module):
    """MaxPool3d(kernel, stride=kernel).  Applied to a non-negative (post-ReLU) input, so the fused
    ReLU+pool kernel is exact here; Features fuses the preceding ReLU into the same launch."""

    def __init__(self, kernel_size, stride=None):
        super().__init__()
        self.kernel_size = ops._triple(kernel_size)
        self.stride = ops._triple(stride if stride is not None else kernel_size)
        assert self.kernel_size == self.stride

    def forward(self, x):
        return ops.relu_maxpool3d(x, self.kernel_size)


class AvgPool3d(nn.Module):
    def __init__(self, kernel_size=2, stride=2):
        super().__init__()
        assert ops._triple(kernel_size) == (2, 2, 2) and ops._triple(stride) == (2, 2, 2)
        self.kernel_size = self.stride = (2, 2, 2)

    def forward(self, x):
        return ops.avgpool3d_2(x)


class InstanceNorm(nn.GroupNorm):
    """nn.GroupNorm(C, C, affine=True) (networks.py:784); only used fused with the ReLU after it."""

    def forward(self, x):
        raise RuntimeError('InstanceNorm is evaluated fused with ReLU by Features.forward')


class Features(nn.Sequential):
    """nn.Sequential with the reference's module indices that fuses [ReLU, MaxPool3d], [InstanceNorm, ReLU, AvgPool3d] and
    [InstanceNorm, ReLU] neighbours into single kernels."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, ReLU) and isinstance(nxt, MaxPool3d):
                x = ops.relu_maxpool3d(x, nxt.kernel_size)
                i += 2
            elif isinstance(m, InstanceNorm) and isinstance(nxt, ReLU) and i + 2 < len(mods) and isinstance(mods[i + 2], AvgPool3d):
                x = ops.instancenorm_relu_avgpool(x, m.weight, m.bias)        # one launch when the net is frozen
                i += 3
            elif isinstance(m, InstanceNorm) and isinstance(nxt, ReLU):
                x = ops.instancenorm_relu(x, m.weight, m.bias)
                i += 2
            else:
                x = m(x)
                i += 1
        return x


class ConvNet3D(nn.Module):
    def __init__(self, channel, num_classes, net_width, net_depth, net_act, net_norm, net_pooling, frames,
                 im_size=(32, 32), dropout_keep_prob=0.5):
        super().__init__()
        self.features, shape_feat = self._make_layers(channel, net_width, net_depth, net_norm, net_act,
                                                      net_pooling, im_size, frames)
        # head (networks.py:733-736): stride-1 average pool, dropout, 1x1x1 conv
        self.avg_pool_kernel = (2, 2, 2) if (im_size[0] > 64) else (2, 1, 1)
        self.avg_pool = nn.AvgPool3d(kernel_size=self.avg_pool_kernel, stride=(1, 1, 1))
        self.dropout = nn.Dropout(dropout_keep_prob)
        self.logit = Conv3d(net_width, num_classes, kernel_size=(1, 1, 1), stride=(1, 1, 1), bias=True)
        self.im_size = tuple(im_size)

    def forward(self, x):
        x = x.permute(0, 2, 1, 3, 4)
        out = self.features(x)
        # the head works on (B,128,<=4,<=2,<=2) tensors: pooling/dropout/max are left to torch
        # (they are double-differentiable there); the 1x1x1 conv is ours.
        out = self.logit(self.dropout(self.avg_pool(out)))
        logits = out.squeeze(3).squeeze(3)
        logits = torch.max(logits, 2)[0]
        return logits

    def embed(self, x):
        x = x.permute(0, 2, 1, 3, 4)
        out = self.features(x)
        out = out.reshape(out.size(0), -1)
        return out

    def _get_activation(self, net_act):
        if net_act == 'relu':
            return ReLU(inplace=True)
        raise NotImplementedError('ConvNet3D (B200): only net_act="relu" is on the distillation path, got %s' % net_act)

    def _get_pooling(self, net_pooling, flag):
        if net_pooling == 'maxpooling':
            return MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2)) if flag == 1 else MaxPool3d(kernel_size=2, stride=2)
        if net_pooling == 'avgpooling':
            return AvgPool3d(kernel_size=2, stride=2)
        if net_pooling == 'none':
            return None
        raise ValueError('unknown net_pooling: %s' % net_pooling)

    def _get_normlayer(self, net_norm, shape_feat):
        if net_norm == 'instancenorm':
            return InstanceNorm(shape_feat[0], shape_feat[0], affine=True)
        if net_norm == 'none':
            return None
        raise NotImplementedError('ConvNet3D (B200): net_norm must be "none" or "instancenorm", got %s' % net_norm)

    def _make_layers(self, channel, net_width, net_depth, net_norm, net_act, net_pooling, im_size, frames):
        layers = []
        in_channels = channel
        if im_size[0] == 28:
            im_size = (32, 32)
        shape_feat = [in_channels, frames, im_size[0], im_size[1]]
        for d in range(net_depth):
            layers += [Conv3d(in_channels, 64 if d == 0 else net_width, kernel_size=KERNEL, padding=PADDING, stride=STRIDE)]
            shape_feat[2] //= 2
            shape_feat[3] //= 2
            shape_feat[0] = 64 if d == 0 else net_width
            if net_norm != 'none':
                layers += [self._get_normlayer(net_norm, shape_feat)]
            layers += [self._get_activation(net_act)]
            in_channels = shape_feat[0]
            if net_pooling != 'none':
                layers += [self._get_pooling(net_pooling, 1 if d == 0 else 0)]
                if d != 0:
                    shape_feat[1] //= 2
                shape_feat[2] //= 2
                shape_feat[3] //= 2
        return Features(*layers), shape_feat
