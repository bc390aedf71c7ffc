"""Oracle restatement of the Distribution-Matching inner loop (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/distill_baseline.py:84-90,334-356 (DM on a leaf synthetic video) and
/root/reference/distill_s2d_ms.py:81-87,393-438 (DM + static/dynamic memory).
Autograd on CPU is used for the backward; the explicit closed-form gradient of the loss is
also provided (dm_loss_grad) because the CUDA path implements that form.

Parity pinning: oracle/make_golden.py -> tests/golden/dm_*.npz.
"""
import numpy as np
import torch

from .composer import compose
from .convnet3d import convnet3d_embed


def sample_real_indices(indices_class, c, n):
    """get_images(c, n) index draw: ``np.random.permutation(indices_class[c])[:n]`` on the numpy
    GLOBAL generator (distill_baseline.py:85, distill_s2d_ms.py:82)."""
    return np.random.permutation(indices_class[c])[:n]


def s2d_sample_indices(num_classes, vpc, spc, coin_dynamic, coin_static):
    """Index arithmetic of distill_s2d_ms.py:402-406.

    ``coin_dynamic`` / ``coin_static`` are the two ``torch.randint(2, (C*vpc,))`` draws, in the
    order the reference makes them (dynamic first :405, static second :406).
    """
    n = num_classes * vpc
    label = torch.arange(num_classes, dtype=torch.long).repeat_interleave(vpc)     # :402
    ran = torch.arange(0, n)                                                       # :403
    idx = ran % vpc                                                                # :404
    dynamic_idx = 2 * idx + coin_dynamic.long()                                    # :405
    static_idx = spc * label + 2 * idx + coin_static.long()                        # :406
    return label, idx, dynamic_idx, static_idx


def dm_loss(emb_real, emb_syn):
    """One class term of distill_baseline.py:351: ``sum((mean(real,0) - mean(syn,0))**2)``."""
    return torch.sum((torch.mean(emb_real, dim=0) - torch.mean(emb_syn, dim=0)) ** 2)


def dm_loss_grad(emb_real, emb_syn):
    """Closed form of d loss / d emb_syn[i] = -(2/ipc) (mu_r - mu_s)  (SURVEY §8 a7)."""
    ipc = emb_syn.shape[0]
    diff = torch.mean(emb_real, dim=0) - torch.mean(emb_syn, dim=0)
    return (-(2.0 / ipc) * diff).unsqueeze(0).expand_as(emb_syn)


def sgd_momentum_step(p, grad, buf, lr, momentum):
    """torch.optim.SGD(momentum=m), dampening 0, no nesterov, no weight decay:
    first step buf = grad, later buf = m*buf + grad; p -= lr*buf for EVERY element
    (dense).  ``buf=None`` means first step.  Returns (p_new, buf_new)."""
    if buf is None:
        buf = grad.clone()
    else:
        buf = momentum * buf + grad
    return p - lr * buf, buf


def dm_baseline_iteration(params, image_syn, real, indices_class, *, ipc, batch_real,
                          net_kw=None, real_idx=None):
    """One iteration body of distill_baseline.py:334-354 up to (and including) backward.

    params       frozen ConvNet3D parameters (dict)
    image_syn    (C*ipc, T, 3, H, W) leaf
    real         (N, T, 3, H, W) CPU tensor; indices_class: list of index lists
    real_idx     optional pre-drawn list of index arrays (else drawn from numpy global RNG)
    Returns dict(loss, grad_image_syn, real_idx, emb_real[list], emb_syn[list]).
    """
    net_kw = net_kw or {}
    num_classes = len(indices_class)
    image_syn = image_syn.detach().clone().requires_grad_(True)
    loss = torch.tensor(0.0, dtype=image_syn.dtype)
    drawn, er, es = [], [], []
    for c in range(num_classes):
        idx = real_idx[c] if real_idx is not None else sample_real_indices(indices_class, c, batch_real)
        drawn.append(np.asarray(idx))
        img_real = real[torch.as_tensor(np.asarray(idx), dtype=torch.long)]
        img_syn = image_syn[c * ipc:(c + 1) * ipc]
        out_real = convnet3d_embed(params, img_real, **net_kw).detach()
        out_syn = convnet3d_embed(params, img_syn, **net_kw)
        loss = loss + dm_loss(out_real, out_syn)
        er.append(out_real)
        es.append(out_syn.detach())
    loss.backward()
    return dict(loss=loss.detach(), grad_image_syn=image_syn.grad.detach(), real_idx=drawn,
                emb_real=er, emb_syn=es)


def dm_s2d_iteration(params, static_syn, dynamic_syn, hal, real, indices_class, *, vpc, spc,
                     batch_real, coin_dynamic, coin_static, train_static=False, net_kw=None,
                     real_idx=None):
    """One iteration body of distill_s2d_ms.py:393-431 up to (and including) backward.

    static_syn (C*spc, 3, H, W); dynamic_syn (C, dpc, T, 1, H, W); hal = dict(encoder.weight/bias).
    Returns loss and the gradients w.r.t. dynamic_syn (dense), the hallucinator and (optionally)
    static_syn, plus all sampled indices and the composed videos.
    """
    net_kw = net_kw or {}
    num_classes = len(indices_class)
    label, idx, dynamic_idx, static_idx = s2d_sample_indices(num_classes, vpc, spc, coin_dynamic, coin_static)
    static_syn = static_syn.detach().clone().requires_grad_(train_static)
    dynamic_syn = dynamic_syn.detach().clone().requires_grad_(True)
    w = hal['encoder.weight'].detach().clone().requires_grad_(True)
    b = hal['encoder.bias'].detach().clone().requires_grad_(True)

    static = static_syn[static_idx]                                    # :409
    dynamic = dynamic_syn[label, dynamic_idx]                          # :410
    image_syn = compose(static, dynamic, w, b)                         # :412

    loss = torch.tensor(0.0, dtype=dynamic_syn.dtype)
    drawn, er, es = [], [], []
    for c in range(num_classes):
        ridx = real_idx[c] if real_idx is not None else sample_real_indices(indices_class, c, batch_real)
        drawn.append(np.asarray(ridx))
        img_real = real[torch.as_tensor(np.asarray(ridx), dtype=torch.long)]
        img_syn = image_syn[c * vpc:(c + 1) * vpc].reshape((vpc,) + tuple(image_syn.shape[1:]))   # :417
        out_real = convnet3d_embed(params, img_real, **net_kw).detach()
        out_syn = convnet3d_embed(params, img_syn, **net_kw)
        loss = loss + dm_loss(out_real, out_syn)
        er.append(out_real)
        es.append(out_syn.detach())
    loss.backward()
    return dict(loss=loss.detach(), grad_dynamic=dynamic_syn.grad.detach(),
                grad_hal_weight=w.grad.detach(), grad_hal_bias=b.grad.detach(),
                grad_static=static_syn.grad.detach() if train_static else None,
                label=label, idx=idx, dynamic_idx=dynamic_idx, static_idx=static_idx,
                real_idx=drawn, image_syn=image_syn.detach(), emb_real=er, emb_syn=es)
