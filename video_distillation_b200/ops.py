"""Python face of the C ABI: raw kernel wrappers and the autograd Functions built on them.

The conv trio (fprop / dgrad / wgrad) is closed under differentiation, and
ReLU+MaxPool routing is linear once its argmax code is fixed, so every Function here has a
backward expressed with the same kernels and is differentiable to any order — this is what the
MTT unroll (torch.autograd.grad(create_graph=True), distill_s2d_ms.py:264) needs in place of
ATen's _convolution_double_backward.  There is no CPU / ATen fallback: tensors must be CUDA.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvGeom, check, lib, ptr, stream


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


def conv_geom(x_shape, w_shape, stride, padding):
    N, Cin, T, H, W = x_shape
    Cout, Cin2, kt, kh, kw = w_shape
    assert Cin == Cin2, f'channel mismatch {Cin} vs {Cin2}'
    st, sh, sw = _triple(stride)
    pt, ph, pw = _triple(padding)
    To, Ho, Wo = (T + 2 * pt - kt) // st + 1, (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    return ConvGeom(N, Cin, T, H, W, Cout, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw)


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError(f'video_distillation_b200: expected float32, got {t.dtype}')
    return t.contiguous()


# ------------------------------------------------------------------ conv backend of the trio
_CONV_BACKEND = 'fp32'


def set_conv_backend(name):
    """'fp32': exact CUDA-core kernels (parity mode).  'tc': the feature-conv geometries of ConvNet3D run as
    tcgen05 GEMMs with fp32 accumulation (tc_trio.py): split fprop (three products per MAC), single-pass bf16 dgrad / wgrad;
    other convolutions stay on the fp32 kernels.  'tc_x3': as 'tc' with dgrad and wgrad split as well (both operands as bf16
    hi + lo pairs, hi*hi + hi*lo + lo*hi accumulated in fp32) — the parity-grade tensor-core trio of the MTT unroll."""
    global _CONV_BACKEND
    if name not in ('fp32', 'tc', 'tc_x3'):
        raise ValueError("conv backend must be 'fp32', 'tc' or 'tc_x3'")
    prev, _CONV_BACKEND = _CONV_BACKEND, name
    return prev


def backend_for_precision(precision):
    """Conv backend of the trio for a driver-level precision string ('fp32' | 'bf16' | 'bf16x3')."""
    return {'fp32': 'fp32', 'bf16': 'tc', 'bf16x3': 'tc_x3'}[precision]


_TC_FPROP_SPLIT = __import__('os').environ.get('VD_TC_FPROP_SPLIT', '1') != '0'
_TC_DGRAD_SPLIT = __import__('os').environ.get('VD_TC_DGRAD_SPLIT', '0') != '0'
_TC_OPS = set(__import__('os').environ.get('VD_TC_OPS', 'fprop,dgrad,wgrad').split(','))     # experiment knob


def _tc_route(x_shape, w_shape, stride, padding, device, op='fprop'):
    if _CONV_BACKEND not in ('tc', 'tc_x3') or op not in _TC_OPS:
        return None, None
    from .tc_trio import trio_for
    return trio_for(tuple(x_shape), tuple(w_shape), _triple(stride), _triple(padding), device)


# ------------------------------------------------------------------ raw kernels
def conv3d_fprop_raw(x, w, bias, stride, padding, fp16_ok=False):
    """fp16_ok: the operands have the range of activations x weights (the network's own forward), so the split fprop of
    conv 1 / conv 2 may run as ONE launch on fp16 hi / lo pairs; cotangent operands (the fprop calls that the double backward
    of dgrad / wgrad makes) keep to bf16 pairs, which have the exponent range of fp32."""
    x, w = _f32c(x), _f32c(w)
    trio, layer = _tc_route(x.shape, w.shape, stride, padding, x.device)
    if trio is not None and x.shape[0] > 0:
        # split-bf16: x = xh + xl, w = wh + wl (bf16 each); xh*wh + xh*wl + xl*wh recovers ~16 mantissa bits.
        # The forward decides ReLU masks / pool argmax and the logits, so it is the precision-critical third of
        # the trio (single-pass bf16 fprop alone costs 5-7 % on MTT's second-order gradients; measured in
        # tests/test_dm_gpu.py::test_mtt_s2d_golden).
        y = trio.fprop(layer, x, w, split=_TC_FPROP_SPLIT, fp16_ok=fp16_ok)
        return y if bias is None else y + bias.view(1, -1, 1, 1, 1)
    g = conv_geom(x.shape, w.shape, stride, padding)
    y = torch.empty(g.N, g.Cout, g.To, g.Ho, g.Wo, dtype=torch.float32, device=x.device)
    if y.numel():
        b = _f32c(bias) if bias is not None else None
        check(lib().vd_conv3d_fprop_f32(ptr(x), ptr(w), ptr(b), ptr(y), ctypes.byref(g), stream()), 'conv3d_fprop')
    return y


def conv3d_dgrad_raw(gy, w, x_shape, stride, padding):
    gy, w = _f32c(gy), _f32c(w)
    g = conv_geom(x_shape, w.shape, stride, padding)
    assert tuple(gy.shape) == (g.N, g.Cout, g.To, g.Ho, g.Wo), (tuple(gy.shape), (g.N, g.Cout, g.To, g.Ho, g.Wo))
    trio, layer = _tc_route(x_shape, w.shape, stride, padding, gy.device, 'dgrad')
    if trio is not None and gy.shape[0] > 0:
        if _TC_DGRAD_SPLIT or _CONV_BACKEND == 'tc_x3':   # same split as fprop: gy = gh + gl, w = wh + wl
            return trio.dgrad_split(layer, gy, w)
        return trio.dgrad(layer, gy, w)
    gx = torch.empty(tuple(x_shape), dtype=torch.float32, device=gy.device)
    if gx.numel():
        check(lib().vd_conv3d_dgrad_f32(ptr(gy), ptr(w), ptr(gx), ctypes.byref(g), stream()), 'conv3d_dgrad')
    return gx


def conv3d_wgrad_raw(x, gy, w_shape, stride, padding, want_bias=False, cache_x=False):
    x, gy = _f32c(x), _f32c(gy)
    g = conv_geom(x.shape, w_shape, stride, padding)
    trio, layer = _tc_route(x.shape, w_shape, stride, padding, x.device, 'wgrad')
    if trio is not None and x.shape[0] > 0:
        if _CONV_BACKEND == 'tc_x3':                      # x = xh + xl, gy = gh + gl: three products, fp32 sums
            gw = trio.wgrad_split(layer, x, gy, cache_x=cache_x)
        else:
            gw = trio.wgrad(layer, x, gy, cache_x=cache_x)
        return (gw, gy.sum(dim=(0, 2, 3, 4))) if want_bias else gw
    gw = torch.zeros(tuple(w_shape), dtype=torch.float32, device=x.device)
    gb = torch.zeros(g.Cout, dtype=torch.float32, device=x.device) if want_bias else None
    if x.numel() and gy.numel():
        check(lib().vd_conv3d_wgrad_f32(ptr(x), ptr(gy), ptr(gw), ptr(gb), ctypes.byref(g), stream()), 'conv3d_wgrad')
    return (gw, gb) if want_bias else gw


# ------------------------------------------------------------------ conv trio as autograd Functions
class _Fprop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding, fp16_ok=False):
        ctx.save_for_backward(x, w)
        ctx.sp = (stride, padding)
        return conv3d_fprop_raw(x, w, None, stride, padding, fp16_ok)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.sp
        gx = _Dgrad.apply(gy, w, tuple(x.shape), stride, padding) if ctx.needs_input_grad[0] else None
        # cache_x: x is the activation saved by this forward — the MTT unroll differentiates it twice (tc_trio._xcol_cache)
        gw = _Wgrad.apply(x, gy, tuple(w.shape), stride, padding, True) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None, None


class _Dgrad(torch.autograd.Function):
    """gx = dgrad(gy, w); linear in both: d/dgy = fprop(c, w), d/dw = wgrad(c, gy)."""
    @staticmethod
    def forward(ctx, gy, w, x_shape, stride, padding):
        ctx.save_for_backward(gy, w)
        ctx.sp = (x_shape, stride, padding)
        return conv3d_dgrad_raw(gy, w, x_shape, stride, padding)

    @staticmethod
    def backward(ctx, c):
        gy, w = ctx.saved_tensors
        x_shape, stride, padding = ctx.sp
        ggy = _Fprop.apply(c, w, stride, padding) if ctx.needs_input_grad[0] else None
        gw = _Wgrad.apply(c, gy, tuple(w.shape), stride, padding) if ctx.needs_input_grad[1] else None
        return ggy, gw, None, None, None


class _Wgrad(torch.autograd.Function):
    """gw = wgrad(x, gy); d/dx = dgrad(gy, c), d/dgy = fprop(x, c)."""
    @staticmethod
    def forward(ctx, x, gy, w_shape, stride, padding, cache_x=False):
        ctx.save_for_backward(x, gy)
        ctx.sp = (w_shape, stride, padding)
        return conv3d_wgrad_raw(x, gy, w_shape, stride, padding, cache_x=cache_x)

    @staticmethod
    def backward(ctx, c):
        x, gy = ctx.saved_tensors
        w_shape, stride, padding = ctx.sp
        gx = _Dgrad.apply(gy, c, tuple(x.shape), stride, padding) if ctx.needs_input_grad[0] else None
        ggy = _Fprop.apply(x, c, stride, padding) if ctx.needs_input_grad[1] else None
        return gx, ggy, None, None, None, None


def conv3d(x, w, bias=None, stride=1, padding=0):
    """F.conv3d replacement (fp32, CUDA cores), differentiable to any order."""
    y = _Fprop.apply(x, w, _triple(stride), _triple(padding), True)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return y


# ------------------------------------------------------------------ ReLU + MaxPool routing
def _pool_out(shape, k):
    N, C, T, H, W = shape
    return (N, C, T // k[0], H // k[1], W // k[2])


class _RouteGather(torch.autograd.Function):
    """y[o] = active(o) ? x[src(o)] : 0 for a fixed routing code (linear in x)."""
    @staticmethod
    def forward(ctx, x, code, k):
        ctx.save_for_backward(code)
        ctx.meta = (tuple(x.shape), k)
        x = _f32c(x)
        N, C, T, H, W = x.shape
        y = torch.empty(_pool_out(x.shape, k), dtype=torch.float32, device=x.device)
        if y.numel():
            check(lib().vd_route_gather_f32(ptr(x), ptr(code), ptr(y), N * C, T, H, W, k[0], k[1], k[2], stream()), 'route_gather')
        return y

    @staticmethod
    def backward(ctx, gy):
        (code,) = ctx.saved_tensors
        shape, k = ctx.meta
        return _RouteScatter.apply(gy, code, shape, k), None, None


class _RouteScatter(torch.autograd.Function):
    """gx[src(o)] = active(o) ? gy[o] : 0 (transpose of the gather)."""
    @staticmethod
    def forward(ctx, gy, code, x_shape, k):
        ctx.save_for_backward(code)
        ctx.meta = (x_shape, k)
        gy = _f32c(gy)
        N, C, T, H, W = x_shape
        gx = torch.empty(x_shape, dtype=torch.float32, device=gy.device)
        if gx.numel():
            check(lib().vd_route_scatter_f32(ptr(gy), ptr(code), ptr(gx), N * C, T, H, W, k[0], k[1], k[2], stream()), 'route_scatter')
        return gx

    @staticmethod
    def backward(ctx, c):
        (code,) = ctx.saved_tensors
        _, k = ctx.meta
        return _RouteGather.apply(c, code, k), None, None, None


def route_scatter_raw(gy, code, x_shape, k):
    """gx[src(o)] = active(o) ? gy[o] : 0 without autograd bookkeeping (backward of the fused tensor-core embed)."""
    gy = _f32c(gy)
    N, C, T, H, W = x_shape
    gx = torch.empty(x_shape, dtype=torch.float32, device=gy.device)
    if gx.numel():
        check(lib().vd_route_scatter_f32(ptr(gy), ptr(code.contiguous()), ptr(gx), N * C, T, H, W, k[0], k[1], k[2], stream()), 'route_scatter')
    return gx


class _ReluMaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        x = _f32c(x)
        N, C, T, H, W = x.shape
        y = torch.empty(_pool_out(x.shape, k), dtype=torch.float32, device=x.device)
        code = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
        if y.numel():
            check(lib().vd_relu_maxpool_fwd_f32(ptr(x), ptr(y), ptr(code), N * C, T, H, W, k[0], k[1], k[2], stream()), 'relu_maxpool_fwd')
        ctx.save_for_backward(code)
        ctx.meta = (tuple(x.shape), k)
        ctx.mark_non_differentiable(code)
        return y, code

    @staticmethod
    def backward(ctx, gy, _gcode):
        (code,) = ctx.saved_tensors
        shape, k = ctx.meta
        return _RouteScatter.apply(gy, code, shape, k), None


def relu_maxpool3d(x, kernel, return_code=False):
    """MaxPool3d(kernel, stride=kernel)(ReLU(x)) fused (networks.py:757,766-770)."""
    y, code = _ReluMaxPool.apply(x, _triple(kernel))
    return (y, code) if return_code else y


def route_with_code(x, code, kernel):
    """Apply a GIVEN routing (ReLU mask + pool argmax) to x — used for routing-conditioned parity."""
    return _RouteGather.apply(x, code, _triple(kernel))


# ------------------------------------------------------------------ instancenorm / avgpool variant
class _InormRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        x = _f32c(x)
        N, C = x.shape[0], x.shape[1]
        S = x.numel() // (N * C)
        y = torch.empty_like(x)
        mean = torch.empty(N * C, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        check(lib().vd_inorm_relu_fwd_f32(ptr(x), ptr(_f32c(gamma)), ptr(_f32c(beta)), ptr(y), ptr(mean), ptr(rstd), N, C, S, stream()), 'inorm_relu_fwd')
        ctx.save_for_backward(x, y, gamma, mean, rstd)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, y, gamma, mean, rstd = ctx.saved_tensors
        N, C = x.shape[0], x.shape[1]
        S = x.numel() // (N * C)
        gx = torch.empty_like(x)
        gg = torch.zeros(C, dtype=torch.float32, device=x.device)
        gb = torch.zeros(C, dtype=torch.float32, device=x.device)
        check(lib().vd_inorm_relu_bwd_f32(ptr(x), ptr(y), ptr(_f32c(gy)), ptr(_f32c(gamma)), ptr(mean), ptr(rstd),
                                          ptr(gx), ptr(gg), ptr(gb), N, C, S, stream()), 'inorm_relu_bwd')
        return gx, gg, gb


def instancenorm_relu(x, gamma, beta):
    """ReLU(GroupNorm(C, C, affine)(x)) fused (networks.py:784,757), first-order differentiable."""
    return _InormRelu.apply(x, gamma, beta)


class _AvgPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        N, C, T, H, W = x.shape
        ctx.shape = tuple(x.shape)
        y = torch.empty(N, C, T // 2, H // 2, W // 2, dtype=torch.float32, device=x.device)
        if y.numel():
            check(lib().vd_avgpool2_fwd_f32(ptr(x), ptr(y), N * C, T, H, W, stream()), 'avgpool2_fwd')
        return y

    @staticmethod
    def backward(ctx, gy):
        return _AvgPool2Bwd.apply(gy, ctx.shape)


class _AvgPool2Bwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, shape):
        gy = _f32c(gy)
        N, C, T, H, W = shape
        gx = torch.empty(shape, dtype=torch.float32, device=gy.device)
        if gx.numel():
            check(lib().vd_avgpool2_bwd_f32(ptr(gy), ptr(gx), N * C, T, H, W, stream()), 'avgpool2_bwd')
        return gx

    @staticmethod
    def backward(ctx, c):
        return _AvgPool2.apply(c), None


def avgpool3d_2(x):
    """AvgPool3d(kernel_size=2, stride=2) (networks.py:772)."""
    return _AvgPool2.apply(x)


def instancenorm_relu_avgpool(x, gamma, beta):
    """AvgPool3d(2)(ReLU(GroupNorm(C, C)(x))) (networks.py:784,757,772).  Frozen network (nothing requires a gradient — the DM /
    evaluation embeds): ONE launch that never writes the normalised activation (vd_inorm_relu_avgpool_fwd_f32); otherwise the
    differentiable pair instancenorm_relu + avgpool3d_2."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or gamma.requires_grad or beta.requires_grad)
    N, C, T, H, W = x.shape
    if needs_grad or W % 2 or (T * H * W) % 4:
        return avgpool3d_2(instancenorm_relu(x, gamma, beta))
    x = _f32c(x)
    y = torch.empty(N, C, T // 2, H // 2, W // 2, dtype=torch.float32, device=x.device)
    if y.numel():
        check(lib().vd_inorm_relu_avgpool_fwd_f32(ptr(x), ptr(_f32c(gamma)), ptr(_f32c(beta)), ptr(y), None, None, N, C, T, H, W, stream()),
              'inorm_relu_avgpool_fwd')
    return y


# ------------------------------------------------------------------ composer
class _Compose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, static_syn, dynamic_syn, weight, bias, static_idx, label, dynamic_idx, unique_rows=False):
        ctx.unique_rows = bool(unique_rows)
        static_syn, dynamic_syn, weight, bias = map(_f32c, (static_syn, dynamic_syn, weight, bias))
        C, dpc, T, one, H, W = dynamic_syn.shape
        assert one == 1 and tuple(weight.shape) == (3, 4, 3, 3, 3) and static_syn.shape[1:] == (3, H, W)
        B = int(label.numel())
        idx = [t.to(torch.int64).contiguous() for t in (static_idx, label, dynamic_idx)]
        out = torch.empty(B, T, 3, H, W, dtype=torch.float32, device=dynamic_syn.device)
        check(lib().vd_compose_fwd_ex_f32(ptr(static_syn), ptr(dynamic_syn), ptr(idx[0]), ptr(idx[1]), ptr(idx[2]),
                                          ptr(weight), ptr(bias), ptr(out), B, T, H, W, dpc,
                                          int(static_syn.shape[0]), int(C * dpc), stream()), 'compose_fwd')
        ctx.save_for_backward(static_syn, dynamic_syn, weight, *idx)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        static_syn, dynamic_syn, weight, sidx, label, didx = ctx.saved_tensors
        C, dpc, T, _, H, W = dynamic_syn.shape
        B = int(label.numel())
        gout = _f32c(gout)
        need_s, need_d, need_w, need_b = ctx.needs_input_grad[:4]
        gd = torch.zeros_like(dynamic_syn)
        gw = torch.zeros_like(weight) if (need_w or need_b) else None
        gb = torch.zeros(3, dtype=torch.float32, device=gout.device) if (need_w or need_b) else None
        gs = torch.zeros_like(static_syn) if need_s else None
        if gw is not None and gs is None:
            # one pass over the video gradient, block sums added in a fixed order (no float atomics): reproducible bit for bit
            scratch = torch.empty(B * ((H + 7) // 8) * 328, dtype=torch.float32, device=gout.device)
            check(lib().vd_compose_bwd_fused_ex_f32(ptr(gout), ptr(static_syn), ptr(dynamic_syn), ptr(sidx), ptr(label), ptr(didx),
                                                    ptr(weight), ptr(gd), ptr(gw), ptr(gb), ptr(scratch), scratch.numel(),
                                                    int(ctx.unique_rows), B, T, H, W, dpc,
                                                    int(static_syn.shape[0]), int(C * dpc), stream()), 'compose_bwd_fused')
        else:
            check(lib().vd_compose_bwd_f32(ptr(gout), ptr(static_syn), ptr(dynamic_syn), ptr(sidx), ptr(label), ptr(didx),
                                           ptr(weight), ptr(gd), ptr(gw), ptr(gb), ptr(gs), B, T, H, W, dpc, stream()), 'compose_bwd')
        return gs, (gd if need_d else None), (gw if need_w else None), (gb if need_b else None), None, None, None, None


def compose(static_syn, dynamic_syn, weight, bias, static_idx, label, dynamic_idx, unique_rows=False):
    """hal(static_syn[static_idx], dynamic_syn[label, dynamic_idx]) in one kernel
    (distill_s2d_ms.py:409-412; utils.py:1186-1197, mode='concat').  unique_rows: the caller guarantees that no two videos of
    the batch select the same dynamic memory (true for the reference's index formula, :405): the backward then writes the
    memory gradient with plain stores instead of atomics."""
    return _Compose.apply(static_syn, dynamic_syn, weight, bias, static_idx, label, dynamic_idx, unique_rows)


# ------------------------------------------------------------------ DM loss / optimiser kernels
def class_mean(emb):
    """(C, n, D) -> (C, D) mean over n (distill_baseline.py:351 torch.mean(output_real, dim=0))."""
    emb = _f32c(emb)
    C, n, D = emb.shape
    out = torch.empty(C, D, dtype=torch.float32, device=emb.device)
    check(lib().vd_class_mean_f32(ptr(emb), ptr(out), C, n, D, stream()), 'class_mean')
    return out


def class_sum_ragged(emb, offsets, C):
    """(n_rows, D) embeddings in class-major order + device int32 offsets (C+1,) -> (C, D) per-class SUMS, rows added in order
    (the per-rank partial sums of the video-sharded multi-GPU DM path)."""
    emb = _f32c(emb)
    D = emb.shape[1]
    out = torch.empty(C, D, dtype=torch.float32, device=emb.device)
    if emb.shape[0] == 0:
        return out.zero_()
    check(lib().vd_class_sum_ragged_f32(ptr(emb), ptr(offsets), ptr(out), C, D, stream()), 'class_sum_ragged')
    return out


class _DMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean_real, emb_syn):
        mean_real, emb_syn = _f32c(mean_real), _f32c(emb_syn)
        C, ns, D = emb_syn.shape
        loss = torch.zeros((), dtype=torch.float32, device=emb_syn.device)
        grad = torch.empty_like(emb_syn)
        class_loss = torch.empty(C, dtype=torch.float32, device=emb_syn.device)       # per-class terms, summed in class order
        check(lib().vd_dm_loss_ex_f32(ptr(mean_real), ptr(emb_syn), ptr(loss), ptr(grad), ptr(class_loss), C, ns, D, 1.0, stream()),
              'dm_loss')
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, grad * g


def dm_loss(mean_real, emb_syn):
    """sum_c || mean_real[c] - mean(emb_syn[c], 0) ||^2 with the gradient wrt emb_syn fused into the
    same launch (distill_baseline.py:351 summed over classes).  mean_real (C,D), emb_syn (C,ns,D)."""
    return _DMLoss.apply(mean_real, emb_syn)


def sgd_momentum_(p, grad, buf, lr, momentum, first_step):
    """In-place dense torch.optim.SGD(momentum) step on raw tensors (no autograd)."""
    assert p.is_contiguous() and grad.is_contiguous() and buf.is_contiguous()
    check(lib().vd_sgd_momentum_f32(ptr(p), ptr(grad), ptr(buf), p.numel(), float(lr), float(momentum), int(bool(first_step)), stream()), 'sgd_momentum')


def sqdist(a, b):
    a, b = _f32c(a), _f32c(b)
    out = torch.zeros((), dtype=torch.float32, device=a.device)
    check(lib().vd_sqdist_f32(ptr(a), ptr(b), ptr(out), a.numel(), stream()), 'sqdist')
    return out
