"""CPU emulation of the tensor-core operand formats (tuning aid, not on the product path).

For one DM class term (n_real frozen real videos + n_syn differentiable synthetic videos through
ConvNet3D.embed, networks.py:747-751) this measures, against an fp64 evaluation of the same
network, what each operand format costs:

  embed relL2, loss rel, d loss / d syn video relL2 (unconditioned) and the same gradient with the
  ReLU masks / pool argmax forced to the fp64 routing (conditioned).

Operand formats of a conv  y = sum_k x_k w_k  (accumulation emulated in fp64):
  fp32        operands as they are
  f16 / bf16  single pass, both operands rounded
  f16x3 / bf16x3   x = xh + xl, w = wh + wl, y = xh wh + xl wh + xh wl
  f16a        activations exact (xh + xl), weights rounded once  (2 passes)
  A+B         frozen real videos in format A, differentiable synthetic videos in format B
The backward (dgrad) uses `bwd` format for (dy, w) the same way.
"""
import argparse
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle.convnet3d import init_convnet3d  # noqa: E402

ST, PD = (1, 2, 2), (1, 3, 3)


def rnd(t, fmt):
    if fmt == 'f16':
        return t.to(torch.float16).to(t.dtype)
    if fmt == 'bf16':
        return t.to(torch.bfloat16).to(t.dtype)
    if fmt == 'tf32':          # round to 10 explicit mantissa bits
        i = t.float().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32).to(t.dtype)
    return t


def parts(t, mode):
    """list of (x_part, w_selector) products for operand pair; returns split of one operand"""
    if mode == 'fp32':
        return [t.float().double()], None
    fmt = mode.rstrip('x3a')
    fmt = 'f16' if mode.startswith('f16') else 'bf16' if mode.startswith('bf16') else 'tf32'
    h = rnd(t.float().double(), fmt)
    l = rnd(t.float().double() - h, fmt)
    return [h, l], fmt


def products(a, b, mode):
    """pairs (a_part, b_part) whose products are summed; a = activation-like, b = weight"""
    if mode == 'fp32':
        return [(a.float().double(), b.float().double())]
    (ah, al), _ = parts(a, mode)
    (bh, bl), _ = parts(b, mode)
    if mode.endswith('x3'):
        return [(ah, bh), (al, bh), (ah, bl)]
    if mode.endswith('a'):
        return [(ah, bh), (al, bh)]
    if mode.endswith('w'):     # weights exact (wh + wl), activations rounded once  (2 passes)
        return [(ah, bh), (ah, bl)]
    return [(ah, bh)]


class EmuConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, fwd, bwd):
        ctx.save_for_backward(x, w)
        ctx.bwd = bwd
        y = None
        for xa, wa in products(x, w, fwd):
            t = F.conv3d(xa, wa, None, ST, PD)
            y = t if y is None else y + t
        return y + b.double().view(1, -1, 1, 1, 1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = None
        for ga, wa in products(gy, w, ctx.bwd):
            t = torch.nn.grad.conv3d_input(x.shape, wa, ga, ST, PD)
            gx = t if gx is None else gx + t
        return gx, None, None, None, None


def pool_forced(y, k, idx, mask):
    """max-pool + ReLU with externally given argmax / mask (routing of the fp64 run)"""
    B, C = y.shape[:2]
    flat = y.reshape(B, C, -1)
    out = torch.gather(flat, 2, idx.reshape(B, C, -1)).reshape(idx.shape)
    return out * mask


def embed(params, x, fwd, bwd, routing=None):
    """x: (B,T,3,H,W).  Returns (emb, routing).  ReLU then MaxPool == MaxPool then ReLU."""
    h = x.permute(0, 2, 1, 3, 4).double()
    rec = []
    for d in range(3):
        w, b = params[f'features.{3 * d}.weight'], params[f'features.{3 * d}.bias']
        y = EmuConv.apply(h, w, b, fwd, bwd)
        k = (1, 2, 2) if d == 0 else (2, 2, 2)
        if routing is None:
            p, idx = F.max_pool3d(y, k, k, return_indices=True)
            mask = (p > 0).double()
            rec.append((idx, mask))
            h = p * mask
        else:
            idx, mask = routing[d]
            h = pool_forced(y, k, idx, mask)
    return h.reshape(h.shape[0], -1), rec


SCALE = [1.0]


def run(params, real, syn, fwd, bwd, routing=None, real_fwd=None):
    syn = syn.clone().requires_grad_(True)
    with torch.no_grad():
        er, _ = embed(params, real, real_fwd or fwd, bwd)
    es, rec = embed(params, syn, fwd, bwd, routing)
    loss = ((er.detach().mean(0) - es.mean(0)) ** 2).sum()
    (loss * SCALE[0]).backward()
    return dict(er=er.detach(), es=es.detach(), loss=loss.detach(), g=syn.grad.detach() / SCALE[0], rec=rec)


def rel(a, b):
    return float((a - b).norm() / b.norm())


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--T', type=int, default=16)
    ap.add_argument('--HW', type=int, default=112)
    ap.add_argument('--n_real', type=int, default=8)
    ap.add_argument('--n_syn', type=int, default=2)
    ap.add_argument('--seeds', type=int, default=2)
    ap.add_argument('--modes', default='fp32,f16,f16a,f16x3,bf16,bf16x3')
    ap.add_argument('--bwd', default=None, help='override the backward format: f16 (with --scale), bf16, bf16x3, f16x3')
    ap.add_argument('--scale', type=float, default=1.0, help='loss scale applied before the backward (power of two)')
    a = ap.parse_args()
    torch.set_num_threads(8)
    for seed in range(a.seeds):
        params = init_convnet3d(100 + seed)
        g = torch.Generator().manual_seed(seed)
        real = torch.randn(a.n_real, a.T, 3, a.HW, a.HW, generator=g)
        syn = torch.randn(a.n_syn, a.T, 3, a.HW, a.HW, generator=g)
        # fp64 truth (operands are the fp32 values)
        ref = run({k: v.double() for k, v in params.items()}, real.double(), syn.double(), 'fp32', 'fp32')
        # fp32 arithmetic (what the reference computes on CPU)
        print(f'seed {seed}: |loss| {float(ref["loss"]):.4e}')
        for mode in a.modes.split(','):
            if mode == 'fp32':
                p32 = params
                s = syn.clone().requires_grad_(True)
                def emb32(x):
                    h = x.permute(0, 2, 1, 3, 4)
                    for d in range(3):
                        h = F.conv3d(h, p32[f'features.{3*d}.weight'], p32[f'features.{3*d}.bias'], ST, PD)
                        k = (1, 2, 2) if d == 0 else (2, 2, 2)
                        h = F.max_pool3d(F.relu(h), k, k)
                    return h.reshape(h.shape[0], -1)
                er = emb32(real).detach(); es = emb32(s)
                loss = ((er.mean(0) - es.mean(0)) ** 2).sum(); loss.backward()
                print(f'  {mode:8s} embed {rel(es.detach().double(), ref["es"]):.2e} real-mean {rel(er.double().mean(0), ref["er"].mean(0)):.2e} '
                      f'loss {abs(float(loss) - float(ref["loss"])) / float(ref["loss"]):.2e} grad(uncond) {rel(s.grad.double(), ref["g"]):.2e}')
                continue
            bwd = mode if mode != 'f16a' else 'f16a'
            bwd = {'f16': 'bf16', 'f16a': 'bf16x3', 'f16x3': 'bf16x3'}.get(mode, mode)
            if a.bwd:
                bwd = a.bwd
            SCALE[0] = a.scale
            real_fwd = None
            if '+' in mode:            # "<real format>+<synthetic format>", e.g. f16+f16x3
                real_fwd, fmode = mode.split('+')
                bwd = a.bwd or 'bf16x3'
            else:
                fmode = mode
            r = run(params, real, syn, fmode, bwd, real_fwd=real_fwd)
            rc = run(params, real, syn, fmode, bwd, routing=ref['rec'], real_fwd=real_fwd)
            # same emulated forward / backward, but with exact operands and forced routing = truth gradient
            print(f'  {mode:8s} embed {rel(r["es"], ref["es"]):.2e} real-mean {rel(r["er"].mean(0), ref["er"].mean(0)):.2e} '
                  f'loss {abs(float(r["loss"]) - float(ref["loss"])) / float(ref["loss"]):.2e} grad(uncond) {rel(r["g"], ref["g"]):.2e} '
                  f'grad(cond) {rel(rc["g"], ref["g"]):.2e}  (bwd {bwd})')
