"""The differentiable conv trio (fprop / dgrad / wgrad) of the three ConvNet3D feature convolutions on tensor
cores: bf16 operands, fp32 accumulation, plain fp32 NCDHW tensors in and out.

``ops.conv3d`` builds the MTT unroll and its double backward from three primitives (``ops._Fprop/_Dgrad/_Wgrad``,
closed under differentiation).  With ``ops.set_conv_backend('tc')`` those primitives route here whenever the call
has the geometry of networks.py:799 (k=(3,7,7), s=(1,2,2), p=(1,3,3)) on a supported video shape; everything else
(the 1x1x1 logit conv, odd shapes) stays on the exact fp32 kernels.  Every launch goes through the C ABI.
"""
import ctypes
import os

import torch

from . import _lib
from .tc import make_plan, tc_supported

_KERNEL, _STRIDE, _PAD = (3, 7, 7), (1, 2, 2), (1, 3, 3)

# im2col cache of the MTT unroll.  The activation x of a feature conv is the column operand of TWO wgrads per iteration: the
# first-order one (inner-loop gradient, autograd.grad with create_graph) and, in the grand backward, the wgrad that
# _Fprop.backward makes for the same saved x.  Between xcol_cache_begin() / xcol_cache_end() (one MTT iteration) the packed
# columns of such an x are kept (1.3 + 1.0 + 0.1 GB per unrolled step and part at 50 videos of 16x3x112x112: HBM is there for it)
# and the second wgrad skips its im2col.  An entry holds a reference to x, so the address it is keyed by cannot be reused.
# Budget: at most 30 % of the device memory, and never below 20 % of it left free — a larger unroll / batch (vpc = 5) simply
# packs its columns again.
_xcol_cache = None
_xcol_cache_bytes = 0


def xcol_cache_begin():
    """Opt-in (VD_XCOL_CACHE=1).  Measured: a cold three-iteration run gains 8 % (0.605 -> 0.557 s, bf16x3), but in a long run
    (bench.py: 13 iterations back to back) the 1 GB cache tensors and the activations of the autograd graph compete for the same
    blocks of the caching allocator and the iteration gets SLOWER (568 -> 646 ms; profiles/r02r_bench_mtt_n1.json) — off by
    default until the cache owns a pre-allocated arena."""
    global _xcol_cache, _xcol_cache_bytes
    if os.environ.get('VD_XCOL_CACHE', '0') == '1':
        _xcol_cache, _xcol_cache_bytes = {}, 0


def xcol_cache_end():
    global _xcol_cache, _xcol_cache_bytes
    _xcol_cache, _xcol_cache_bytes = None, 0


def _xcol_cache_admits(nbytes, device):
    total = torch.cuda.get_device_properties(device).total_memory
    if _xcol_cache_bytes + nbytes > 0.30 * total:
        return False
    free = torch.cuda.mem_get_info(device)[0] + torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return free - nbytes > 0.20 * total


class TcTrio:
    def __init__(self, T, H, W, device):
        self.plan = make_plan(T, H, W)
        self.device = torch.device(device)
        p = self.plan
        # (Cin, Cout, input extent) per layer
        self.layers = [
            (3, 64, (p.T, p.H, p.W)),
            (64, 128, (p.T1p, p.H1p, p.W1p)),
            (128, 128, (p.T2p, p.H2p, p.W2p)),
        ]
        self._ws = {}
        self.direct_dgrad1 = True
        self.direct_dgrad0 = True
        self.col_fp32 = True
        self.fused_split_fprop = True     # layers 1 / 2: split fprop as one launch on the split-fp16 tables
        self._last_xcol = {}

    # ------------------------------------------------------------------ helpers
    def layer_of(self, cin, cout, extent):
        for l, (ci, co, ext) in enumerate(self.layers):
            if ci == cin and co == cout and tuple(ext) == tuple(extent):
                return l
        return None

    def _buf(self, name, nbytes, zero=False):
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes:
            t = (torch.zeros if zero else torch.empty)(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[name] = t
        return t

    def _pack_input(self, layer, x, part=0):
        p, lib, B = self.plan, _lib.lib(), int(x.shape[0])
        if layer == 0:
            buf = self._buf('x0', B * p.x0_bytes_per_video)
            _lib.check(lib.vd_tc_pack_video_ncdhw(_lib.ptr(x), _lib.ptr(buf), ctypes.byref(p), B, part, _lib.stream()), 'vd_tc_pack_video_ncdhw')
        elif layer == 1:
            buf = self._buf('a1', B * p.a1_bytes_per_video)
            _lib.check(lib.vd_tc_pack_act(1, _lib.ptr(x), _lib.ptr(buf), ctypes.byref(p), B, part, _lib.stream()), 'vd_tc_pack_act(1)')
        else:
            buf = self._buf('a2', (B + 3) // 4 * 4 * p.a2_bytes_per_video, zero=True)
            _lib.check(lib.vd_tc_pack_act(2, _lib.ptr(x), _lib.ptr(buf), ctypes.byref(p), B, part, _lib.stream()), 'vd_tc_pack_act(2)')
        return buf

    def _pack_weight(self, layer, w, part, name):
        p, lib = self.plan, _lib.lib()
        wimg = self._buf(name, (p.w0_bytes, p.w1_bytes, p.w2_bytes)[layer])
        ws = [None, None, None]
        imgs = [None, None, None]
        ws[layer], imgs[layer] = w, wimg
        _lib.check(lib.vd_tc_pack_weights_part(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]), _lib.ptr(imgs[0]), _lib.ptr(imgs[1]),
                                               _lib.ptr(imgs[2]), part, _lib.stream()), 'vd_tc_pack_weights_part')
        return wimg

    def _conv(self, layer, src, wimg, y, B, accumulate):
        _lib.check(_lib.lib().vd_tc_conv_layer(layer, _lib.ptr(src), _lib.ptr(wimg), None, _lib.ptr(y), None, 0, ctypes.byref(self.plan),
                                               None, B, 3 if accumulate else 2, _lib.stream()), f'vd_tc_conv_layer({layer}, plain)')

    # ------------------------------------------------------------------ the trio
    def fprop(self, layer, x, w, split=False, fp16_ok=False):
        """y = conv3d(x, w).  split: x = xh + xl, w = wh + wl (bf16 parts made inside the packers);
        y = xh*wh + xh*wl + xl*wh accumulated in fp32 by the plain epilogue (three launches, no temporaries).
        fp16_ok (activation x weight ranges): layers 1 / 2 run the three products in ONE launch on fp16 pairs instead —
        not for cotangent operands (1e-5 and below: fp16 keeps no low part there)."""
        p, B = self.plan, int(x.shape[0])
        cin, cout, _ = self.layers[layer]
        out_ext = [(p.T1, p.H1, p.W1), (p.T2, p.H2, p.W2), (p.T3, p.H3, p.W3)][layer]
        y = torch.empty(B, cout, *out_ext, dtype=torch.float32, device=self.device)
        if split and fp16_ok and layer in (1, 2) and self.fused_split_fprop:
            # ONE launch on the split-fp16 tables (fp16 hi / lo pairs, three products accumulated in TMEM)
            lib, plan, st = _lib.lib(), ctypes.byref(p), _lib.stream()
            sz = (ctypes.c_int64 * 6)()
            _lib.check(lib.vd_tc_x3_sizes(plan, sz), 'vd_tc_x3_sizes')
            B4 = (B + 3) // 4 * 4 if layer == 2 else B
            src = self._buf(f'a{layer}s', B4 * int(sz[layer]), zero=(layer == 2))
            wimg = self._buf(f'w{layer}s', int(sz[3 + layer]))
            ws, imgs = [None, None, None], [None, None, None]
            ws[layer], imgs[layer] = w, wimg
            _lib.check(lib.vd_tc_x3_pack_weights(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]), _lib.ptr(imgs[0]), _lib.ptr(imgs[1]),
                                                 _lib.ptr(imgs[2]), st), 'vd_tc_x3_pack_weights')
            _lib.check(lib.vd_tc_x3_pack_act(layer, _lib.ptr(x), _lib.ptr(src), plan, B, st), 'vd_tc_x3_pack_act')
            _lib.check(lib.vd_tc_x3_conv_plain(layer, _lib.ptr(src), _lib.ptr(wimg), None, _lib.ptr(y), plan, B, st), 'vd_tc_x3_conv_plain')
            return y
        wh = self._pack_weight(layer, w, 0, f'w{layer}')
        src = self._pack_input(layer, x, 0)
        self._conv(layer, src, wh, y, B, False)
        if split:
            wl = self._pack_weight(layer, w, 1, f'wl{layer}')
            self._conv(layer, src, wl, y, B, True)
            src = self._pack_input(layer, x, 1)
            self._conv(layer, src, wh, y, B, True)
        return y

    def pack_dgrad_weights(self, layer, w):
        """fp32 OIDHW weights of feature conv `layer` -> the weight image(s) of its dgrad kernel (fresh tensors, so that the
        images of several weight parts can be cached side by side)."""
        p, lib = self.plan, _lib.lib()
        u8 = dict(dtype=torch.uint8, device=self.device)
        plan, st = ctypes.byref(p), _lib.stream()
        if layer == 0 and self.direct_dgrad0:
            sz = (ctypes.c_int64 * 2)()
            _lib.check(lib.vd_tc_dgrad0_sizes(plan, sz), 'vd_tc_dgrad0_sizes')
            wimg = torch.empty(int(sz[1]), **u8)
            _lib.check(lib.vd_tc_pack_dgrad0_weights(_lib.ptr(w), _lib.ptr(wimg), st), 'vd_tc_pack_dgrad0_weights')
            return (wimg,)
        if layer == 1 and self.direct_dgrad1:
            sz = (ctypes.c_int64 * 3)()
            _lib.check(lib.vd_tc_dgrad1_sizes(plan, sz), 'vd_tc_dgrad1_sizes')
            w0, w1 = torch.empty(int(sz[1]), **u8), torch.empty(int(sz[2]), **u8)
            _lib.check(lib.vd_tc_pack_dgrad1_weights(_lib.ptr(w), _lib.ptr(w0), _lib.ptr(w1), plan, st), 'vd_tc_pack_dgrad1_weights')
            return (w0, w1)
        wt = torch.empty(int((p.wt0_bytes, p.wt1_bytes, p.wt2_bytes)[layer]), **u8)
        ws = [None, None, None]
        imgs = [None, None, None]
        ws[layer], imgs[layer] = w, wt
        _lib.check(lib.vd_tc_pack_weights_bwd(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]), _lib.ptr(imgs[0]), _lib.ptr(imgs[1]),
                                              _lib.ptr(imgs[2]), st), 'vd_tc_pack_weights_bwd')
        return (wt,)

    def dgrad(self, layer, gy, w, part=0, wpack=None, out=None, accumulate=False, ncdhw=True, packed=False):
        """gx = dgrad(part(gy), w) of feature conv `layer`.  part: 0 = bf16(gy), 1 = bf16(gy - bf16(gy)) (made inside the packer);
        wpack: cached pack_dgrad_weights(layer, w); out / accumulate: write (or add) into an existing fp32 tensor — the three
        passes of a split-bf16 dgrad then share one tensor; ncdhw=False (layer 0 only): out is (B,T,3,H,W); packed: the operand
        of this (gy, part) is still in the workspace from the previous call (same layer, same batch) and is not packed again."""
        p, lib, B = self.plan, _lib.lib(), int(gy.shape[0])
        cin, cout, ext = self.layers[layer]
        if wpack is None:
            wpack = self.pack_dgrad_weights(layer, w)
        plan, st = ctypes.byref(p), _lib.stream()
        assert ncdhw or layer == 0
        if layer == 0 and self.direct_dgrad0:
            # column-free dgrad of conv 0 (tc_layout.h: Dg0Geo): pixels on M, fp32 accumulators to (B,3,T,H,W) / (B,T,3,H,W)
            sz = (ctypes.c_int64 * 2)()
            _lib.check(lib.vd_tc_dgrad0_sizes(plan, sz), 'vd_tc_dgrad0_sizes')
            dyp = self._buf('dyp0', B * sz[0])
            if not packed:
                _lib.check(lib.vd_tc_pack_dyp0_part(_lib.ptr(gy), _lib.ptr(dyp), plan, B, int(part), st), 'vd_tc_pack_dyp0_part')
            if out is None:
                assert not accumulate
                out = torch.empty((B, cin, *ext) if ncdhw else (B, ext[0], cin, ext[1], ext[2]), dtype=torch.float32, device=self.device)
            _lib.check(lib.vd_tc_dgrad0_ex(_lib.ptr(dyp), _lib.ptr(wpack[0]), _lib.ptr(out), plan, B, 1 if ncdhw else 0, int(accumulate), st),
                       'vd_tc_dgrad0_ex')
            return out
        if layer == 1 and self.direct_dgrad1:
            # column-free dgrad of conv 1 (tc_layout.h: Dg1Geo): fp32 accumulators go straight to the NCDHW gradient
            sz = (ctypes.c_int64 * 3)()
            _lib.check(lib.vd_tc_dgrad1_sizes(plan, sz), 'vd_tc_dgrad1_sizes')
            dyp = self._buf('dyp1', B * sz[0])
            if not packed:
                _lib.check(lib.vd_tc_pack_dyp1_part(_lib.ptr(gy), _lib.ptr(dyp), plan, B, int(part), st), 'vd_tc_pack_dyp1_part')
            if out is None:
                assert not accumulate
                out = torch.empty(B, cin, *ext, dtype=torch.float32, device=self.device)
            _lib.check(lib.vd_tc_dgrad1_plain(_lib.ptr(dyp), _lib.ptr(wpack[0]), _lib.ptr(wpack[1]), _lib.ptr(out), plan, B, int(accumulate), st),
                       'vd_tc_dgrad1_plain')
            return out
        dy = self._buf('dy', B * (p.dy0_bytes_per_video, p.dy1_bytes_per_video, p.dy2_bytes_per_video)[layer])
        f32 = 1 if self.col_fp32 else 0      # fp32 column buffers: no bf16 rounding between the GEMM and the tap sum
        col = self._buf('col', B * (p.col0_bytes_per_video, p.col1_bytes_per_video, p.col2_bytes_per_video)[layer] * (2 if f32 else 1))
        if not packed:
            _lib.check(lib.vd_tc_pack_dy_part(layer, _lib.ptr(gy), _lib.ptr(dy), plan, B, int(part), st), 'vd_tc_pack_dy_part')
        _lib.check(lib.vd_tc_bwd_gemm_ex(layer, _lib.ptr(dy), _lib.ptr(wpack[0]), _lib.ptr(col), plan, B, f32, st), 'vd_tc_bwd_gemm_ex')
        gx = torch.empty(B, cin, *ext, dtype=torch.float32, device=self.device)
        _lib.check(lib.vd_tc_bwd_col2im_plain(layer, _lib.ptr(col), _lib.ptr(gx), plan, B, f32, st), 'vd_tc_bwd_col2im_plain')
        if out is not None:
            if accumulate:
                out += gx
            else:
                out.copy_(gx)
            return out
        return gx

    def dgrad_split(self, layer, gy, w):
        """gx = dgrad(gy, w) on bf16 hi / lo pairs: gh*wh + gh*wl + gl*wh summed in fp32 in one tensor.  The gradient parts are made
        inside the packers, gh is packed once for both weight parts."""
        wh = w.to(torch.bfloat16).float()
        pk_h, pk_l = self.pack_dgrad_weights(layer, wh), self.pack_dgrad_weights(layer, w - wh)
        gx = self.dgrad(layer, gy, None, part=0, wpack=pk_h)
        if layer == 1 and self.direct_dgrad1:
            # separate outputs + elementwise adds: the read-modify-write of conv 1's plain epilogue costs 0.52 ms per launch
            # against 0.17 ms for the plain store (+ 0.08 ms per add), measured at 50 videos
            gx += self.dgrad(layer, gy, None, part=0, wpack=pk_l, packed=True)
            gx += self.dgrad(layer, gy, None, part=1, wpack=pk_h)
            return gx
        self.dgrad(layer, gy, None, part=0, wpack=pk_l, out=gx, accumulate=True, packed=True)
        self.dgrad(layer, gy, None, part=1, wpack=pk_h, out=gx, accumulate=True)
        return gx

    def wgrad_split(self, layer, x, gy, cache_x=False):
        """gw = wgrad(x, gy) on bf16 hi / lo pairs: xl*gh + xh*gh + xh*gl (parts made inside the packers; in this order the
        im2col of xh and the image of gh are each packed once for two GEMMs)."""
        gw = self.wgrad(layer, x, gy, parts=(1, 0), cache_x=cache_x)
        gw += self.wgrad(layer, x, None, parts=(0, 0), cache_x=cache_x)
        gw += self.wgrad(layer, None, gy, parts=(0, 1))
        return gw

    def wgrad(self, layer, x, gy, parts=None, cache_x=False):
        """parts = (x_part, gy_part), 0 = value / 1 = bf16 residual; x or gy None: that operand is still packed in the workspace
        from the previous call of the same layer and batch.  cache_x: x is an activation saved by the forward (see _xcol_cache)."""
        return self._wgrad_parts(layer, x, gy, parts if parts is not None else (0, 0), cache_x)

    def _xcol_for(self, layer, x, part, nbytes, cache_x):
        """(column buffer for (x, part), True when it still has to be packed)."""
        if x is None:
            return self._last_xcol[layer], False
        if cache_x and _xcol_cache is not None:
            key = (id(self), layer, int(part), x.data_ptr(), x._version, tuple(x.shape))
            hit = _xcol_cache.get(key)
            if hit is not None:
                return hit[1], False
            if _xcol_cache_admits(int(nbytes), self.device):
                global _xcol_cache_bytes
                buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)   # every byte is written by the packer
                _xcol_cache[key] = (x, buf)
                _xcol_cache_bytes += int(nbytes)
                return buf, True
        # per-layer workspaces: an operand kept for the next call must not be overwritten by another layer's wgrad in between
        return self._buf(f'xcol{layer}', nbytes), True

    def _wgrad_parts(self, layer, x, gy, parts, cache_x=False):
        p, lib = self.plan, _lib.lib()
        B = int((x if x is not None else gy).shape[0])
        cin, cout, _ = self.layers[layer]
        sizes = (ctypes.c_int64 * 6)()
        _lib.check(lib.vd_tc_wgrad_plan(layer, ctypes.byref(p), B, sizes), 'vd_tc_wgrad_plan')
        xcol, pack_x = self._xcol_for(layer, x, parts[0], sizes[3], cache_x)
        self._last_xcol[layer] = xcol
        gyimg = self._buf(f'gyimg{layer}', sizes[4])
        raw = self._buf('wraw', sizes[5])
        plan, st = ctypes.byref(p), _lib.stream()
        if pack_x or gy is not None:
            _lib.check(lib.vd_tc_wgrad_pack_parts(layer, _lib.ptr(x) if pack_x else None, int(parts[0]), _lib.ptr(gy), int(parts[1]),
                                                  _lib.ptr(xcol), _lib.ptr(gyimg), plan, B, st), 'vd_tc_wgrad_pack_parts')
        _lib.check(lib.vd_tc_wgrad_gemm(layer, _lib.ptr(xcol), _lib.ptr(gyimg), _lib.ptr(raw), plan, B, st), 'vd_tc_wgrad_gemm')
        gw = torch.empty(cout, cin, 3, 7, 7, dtype=torch.float32, device=self.device)
        _lib.check(lib.vd_tc_wgrad_reduce(layer, _lib.ptr(raw), _lib.ptr(gw), plan, B, st), 'vd_tc_wgrad_reduce')
        return gw


_trios = {}


def trio_for(x_shape, w_shape, stride, padding, device):
    """(TcTrio, layer) when the conv (x_shape NCDHW, w_shape OIDHW) is one of the three feature convolutions of a
    supported video shape, else (None, None)."""
    if tuple(w_shape[2:]) != _KERNEL or tuple(stride) != _STRIDE or tuple(padding) != _PAD:
        return None, None
    N, cin, T, H, W = x_shape
    cout = w_shape[0]
    # recover the video shape of the plan this layer belongs to
    if cin == 3:
        vid = (T, H, W)
    elif cin == 64:
        vid = (T, H * 4, W * 4)
    elif cin == 128:
        vid = (T * 2, H * 16, W * 16)
    else:
        return None, None
    cands = [vid] if cin != 128 else [vid, (T * 2, 112, 112), (T * 2, 64, 64)]
    for (t, h, w) in cands:
        if not tc_supported(t, h, w):
            continue
        key = (t, h, w, str(device))
        trio = _trios.get(key)
        if trio is None:
            trio = _trios[key] = TcTrio(t, h, w, device)
        layer = trio.layer_of(cin, cout, (T, H, W))
        if layer is not None:
            return trio, layer
    return None, None
