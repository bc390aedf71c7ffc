"""Oracle restatement of the MTT unrolled-student inner loop (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/distill_baseline.py:213-272 (baseline MTT) and
/root/reference/distill_s2d_ms.py:220-292 (MTT + static/dynamic memory), with the student
being ReparamModule(ConvNet3D) (/root/reference/reparam_module.py:144-159) in train mode
(Dropout(0.5) live, networks.py:735,741).  Second-order autograd on CPU.

All random draws (randperm, the two coin flips, the dropout masks, the start epoch) are
explicit inputs so that a CUDA run that drew them on the device can be replayed here.

Parity pinning: oracle/make_golden.py -> tests/golden/mtt_*.npz.
"""
import torch
import torch.nn.functional as F

from .composer import compose
from .convnet3d import convnet3d_forward, unflatten_params


def mtt_sample_step_indices(perm, vpc, spc, coin_dynamic, coin_static):
    """distill_s2d_ms.py:242-246 for one inner step whose chunk of the randperm is ``perm``."""
    label = perm // vpc
    idx = perm % vpc
    dynamic_idx = 2 * idx + coin_dynamic.long()
    static_idx = spc * label + 2 * idx + coin_static.long()
    return label, idx, dynamic_idx, static_idx


def _unroll(theta0, target, like, step_inputs, syn_lr, im_size, net_kw):
    """theta_{k+1} = theta_k - syn_lr * grad_theta CE(f(x_k; theta_k), y_k) with create_graph
    (distill_s2d_ms.py:261-266), then the normalised parameter distance (:268-282)."""
    student = [theta0.detach().clone().requires_grad_(True)]
    ce_losses = []
    for (x, y, mask) in step_inputs:
        p = unflatten_params(student[-1], like)
        logits = convnet3d_forward(p, x, im_size, dropout_mask=mask, **net_kw)
        ce = F.cross_entropy(logits, y)
        grad = torch.autograd.grad(ce, student[-1], create_graph=True)[0]
        student.append(student[-1] - syn_lr * grad)
        ce_losses.append(ce.detach())
    num_params = theta0.numel()
    param_loss = F.mse_loss(student[-1], target, reduction='sum') / num_params
    param_dist = F.mse_loss(theta0.detach(), target, reduction='sum') / num_params
    return param_loss / param_dist, param_loss.detach(), param_dist.detach(), ce_losses, student[-1].detach()


def mtt_s2d_iteration(theta0, target, like, static_syn, dynamic_syn, hal, syn_lr, *, vpc, spc,
                      perms, coins_dynamic, coins_static, dropout_masks, im_size,
                      train_static=False, train_lr=True, net_kw=None):
    """One iteration body of distill_s2d_ms.py:220-292 (up to and including backward).

    theta0 / target : flat expert parameters at start_epoch / start_epoch+expert_epochs
    like            : dict name -> tensor giving the parameter shapes / order
    perms[k], coins_*[k], dropout_masks[k] : the draws of inner step k
    """
    net_kw = net_kw or {}
    static_syn = static_syn.detach().clone().requires_grad_(train_static)
    dynamic_syn = dynamic_syn.detach().clone().requires_grad_(True)
    w = hal['encoder.weight'].detach().clone().requires_grad_(True)
    b = hal['encoder.bias'].detach().clone().requires_grad_(True)
    syn_lr = syn_lr.detach().clone().requires_grad_(train_lr)
    steps = []
    for perm, cd, cs, mask in zip(perms, coins_dynamic, coins_static, dropout_masks):
        label, idx, dynamic_idx, static_idx = mtt_sample_step_indices(perm, vpc, spc, cd, cs)
        x = compose(static_syn[static_idx], dynamic_syn[label, dynamic_idx], w, b)    # :249-253
        steps.append((x, label.long(), mask))
    grand, ploss, pdist, ces, theta_n = _unroll(theta0, target, like, steps, syn_lr, im_size, net_kw)
    grand.backward()
    return dict(grand_loss=grand.detach(), param_loss=ploss, param_dist=pdist, ce_losses=ces,
                theta_final=theta_n, grad_dynamic=dynamic_syn.grad.detach(),
                grad_hal_weight=w.grad.detach(), grad_hal_bias=b.grad.detach(),
                grad_static=static_syn.grad.detach() if train_static else None,
                grad_syn_lr=syn_lr.grad.detach() if train_lr else None)


def mtt_baseline_iteration(theta0, target, like, image_syn, label_syn, syn_lr, *, perms,
                           dropout_masks, im_size, train_lr=True, net_kw=None):
    """One iteration body of distill_baseline.py:213-272 (leaf synthetic videos)."""
    net_kw = net_kw or {}
    image_syn = image_syn.detach().clone().requires_grad_(True)
    syn_lr = syn_lr.detach().clone().requires_grad_(train_lr)
    steps = [(image_syn[perm], label_syn[perm], mask) for perm, mask in zip(perms, dropout_masks)]
    grand, ploss, pdist, ces, theta_n = _unroll(theta0, target, like, steps, syn_lr, im_size, net_kw)
    grand.backward()
    return dict(grand_loss=grand.detach(), param_loss=ploss, param_dist=pdist, ce_losses=ces,
                theta_final=theta_n, grad_image_syn=image_syn.grad.detach(),
                grad_syn_lr=syn_lr.grad.detach() if train_lr else None)
