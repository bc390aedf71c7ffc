"""Host driver of the tensor-core (tcgen05) ConvNet3D feature pipeline.

``TcConvNet3D`` owns the packed operands of one frozen ConvNet3D (weights as UMMA images, the
per-layer activation buffers) and runs conv0 -> conv1 -> conv2 with their fused
bias+ReLU+MaxPool epilogues, i.e. ``ConvNet3D.embed`` of the reference
(networks.py:747-751 with get_network's none/maxpooling setting, utils.py:608-609) in bf16
operands / fp32 accumulation.  Every launch goes through the C ABI (include/vd_b200.h).
"""
import ctypes

import torch

from . import _lib


def make_plan(T, H, W):
    plan = _lib.TcPlan()
    _lib.check(_lib.lib().vd_tc_plan_make(ctypes.byref(plan), int(T), int(H), int(W)), 'vd_tc_plan_make')
    return plan


def tc_supported(T, H, W):
    # mirrors geo_supported (csrc/tc_layout.h): 112x112 with T in {4, 8, 12, 16} (conv 2 tile <= 128 columns), 64x64 with
    # T in {8, 16, 24, 32} (conv 2 tile = 2T columns, a multiple of 16)
    if H != W or H not in (112, 64):
        return False
    return (T % 4 == 0 and 4 <= T <= 16) if H == 112 else (T % 8 == 0 and 8 <= T <= 32)


class TcConvNet3D:
    """Tensor-core embed of a ConvNet3D whose feature weights are given as fp32 tensors.

    split=False: bf16 operands and bf16 activations between the layers (single pass, throughput mode).
    split=True : "f16x3" — every operand is an fp16 hi/lo pair and every product xh*wh + xl*wh + xh*wl in the same fp32
                 TMEM accumulator (csrc/tc_layout.h: SGeo): embeddings within ~3e-6 of an fp32 evaluation, ReLU / MaxPool
                 routing decided on fp32-equivalent sums; the backward runs the dgrads as split-bf16 (three passes each).
    real_products=2 (split only): the FROZEN real videos (``embed_resident``, the real part of ``embed_joint``, ``embed`` without
                 codes when ``frozen=True``) run in the two-product mode xh*wh + xh*wl — weights exact, each activation
                 rounded once to fp16 (independent from video to video: it averages out of the class means), 2/3 of the MMAs
                 and half of the operand bytes.  Differentiable videos always take the full three products."""

    def __init__(self, T, H, W, device, max_batch=128, split=False, real_products=3):
        self.plan = make_plan(T, H, W)
        self.T, self.H, self.W = T, H, W
        self.device = torch.device(device)
        self.max_batch = int(max_batch)
        self.split = bool(split)
        if real_products not in (2, 3):
            raise ValueError('real_products must be 2 or 3')
        self.real_products = int(real_products) if self.split else 3
        p = self.plan
        u8 = dict(dtype=torch.uint8, device=self.device)
        if self.split:
            sz = (ctypes.c_int64 * 6)()
            _lib.check(_lib.lib().vd_tc_x3_sizes(ctypes.byref(p), sz), 'vd_tc_x3_sizes')
            self.x0_per, self.a1_per, self.a2_per = int(sz[0]), int(sz[1]), int(sz[2])
            self.x0h_per = int(p.x0_bytes_per_video)         # hi-only operand of the two-product mode (X0 layout, fp16 values)
            wb = (int(sz[3]), int(sz[4]), int(sz[5]))
        else:
            self.x0_per, self.a1_per, self.a2_per = int(p.x0_bytes_per_video), int(p.a1_bytes_per_video), int(p.a2_bytes_per_video)
            wb = (int(p.w0_bytes), int(p.w1_bytes), int(p.w2_bytes))
        self.w0 = torch.empty(wb[0], **u8)
        self.w1 = torch.empty(wb[1], **u8)
        self.w2 = torch.empty(wb[2], **u8)
        self._trio = None
        self.codes_override = None   # (c0, c1, c2): embed_backward uses these ReLU / MaxPool routing codes instead of its own
        self.wt0 = torch.empty(p.wt0_bytes, **u8)       # transposed images for the backward column GEMMs
        self.wt1 = torch.empty(p.wt1_bytes, **u8)
        self.wt2 = torch.empty(p.wt2_bytes, **u8)
        self._bwd_ready = False
        self._fp32_w = None
        self._split_bwd_w = None    # per layer: (dgrad weight images of wh, of wl), packed on the first split backward
        self.bwd_chunk = 64
        sz = (ctypes.c_int64 * 3)()
        _lib.check(_lib.lib().vd_tc_dgrad1_sizes(ctypes.byref(p), sz), 'vd_tc_dgrad1_sizes')
        self.dyp1_bytes_per_video = int(sz[0])
        self.dg1_w = (torch.empty(int(sz[1]), **u8), torch.empty(int(sz[2]), **u8))   # direct conv-1 dgrad weight images
        self.direct_dgrad1 = True          # False: column GEMM + col2im for conv 1 as for conv 2
        sz0 = (ctypes.c_int64 * 2)()
        _lib.check(_lib.lib().vd_tc_dgrad0_sizes(ctypes.byref(p), sz0), 'vd_tc_dgrad0_sizes')
        self.dyp0_bytes_per_video = int(sz0[0])
        self.dg0_w = torch.empty(int(sz0[1]), **u8)                                   # direct conv-0 dgrad weight image
        self.direct_dgrad0 = True          # needs direct_dgrad1 (its epilogue writes the padded planar dY of conv 0)
        self._ws = {}
        self.b0 = self.b1 = self.b2 = None
        self._x0 = None
        self._a1 = None
        self._a2 = None
        self.embed_dim = int(p.embed_dim)
        self.timing = None          # set to [] to collect (layer, B, start_event, end_event, products per MAC) per launch

    # ---------------------------------------------------------------- operands
    def load_weights(self, w0, b0, w1, b1, w2, b2):
        """fp32 OIDHW weights/biases of features.{0,3,6} -> UMMA images (3 small pack launches)."""
        ws = [t.detach().contiguous().float() for t in (w0, w1, w2)]
        pack = _lib.lib().vd_tc_x3_pack_weights if self.split else _lib.lib().vd_tc_pack_weights
        _lib.check(pack(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]), _lib.ptr(self.w0), _lib.ptr(self.w1), _lib.ptr(self.w2),
                        _lib.stream()), 'vd_tc_pack_weights')
        self.b0, self.b1, self.b2 = (t.detach().contiguous().float() for t in (b0, b1, b2))
        self._fp32_w = ws
        self._bwd_ready = False
        self._split_bwd_w = None
        return self

    def _prepare_bwd(self):
        if not self._bwd_ready:
            ws = self._fp32_w
            _lib.check(_lib.lib().vd_tc_pack_weights_bwd(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]),
                                                         _lib.ptr(self.wt0), _lib.ptr(self.wt1), _lib.ptr(self.wt2),
                                                         _lib.stream()), 'vd_tc_pack_weights_bwd')
            _lib.check(_lib.lib().vd_tc_pack_dgrad1_weights(_lib.ptr(ws[1]), _lib.ptr(self.dg1_w[0]), _lib.ptr(self.dg1_w[1]),
                                                            ctypes.byref(self.plan), _lib.stream()), 'vd_tc_pack_dgrad1_weights')
            _lib.check(_lib.lib().vd_tc_pack_dgrad0_weights(_lib.ptr(ws[0]), _lib.ptr(self.dg0_w), _lib.stream()), 'vd_tc_pack_dgrad0_weights')
            self._bwd_ready = True

    def _workspace(self, name, nbytes, dtype=torch.uint8, zero=False):
        t = self._ws.get(name)
        n = nbytes // torch.empty((), dtype=dtype).element_size()
        if t is None or t.numel() < n:
            t = (torch.zeros if zero else torch.empty)(n, dtype=dtype, device=self.device)
            self._ws[name] = t
        return t

    def embed_backward(self, g_emb, codes):
        """Gradient of ``embed`` w.r.t. its input videos for the routing recorded in ``codes``
        (weights frozen): three column GEMMs on tensor cores + col2im/routing gathers."""
        if self.codes_override is not None:
            codes = self.codes_override           # routing-conditioned parity (tests): backward along a GIVEN routing
        if self.split:
            return self._embed_backward_split(g_emb, codes)
        self._prepare_bwd()
        p, lib = self.plan, _lib.lib()
        g_emb = g_emb.contiguous().float()
        B = g_emb.shape[0]
        c0, c1, c2 = codes
        dvideo = torch.empty(B, self.T, 3, self.H, self.W, dtype=torch.float32, device=self.device)
        for s in range(0, B, self.bwd_chunk):
            e = min(B, s + self.bwd_chunk)
            n = e - s
            dy2 = self._workspace('dy2', n * p.dy2_bytes_per_video)
            dy1 = self._workspace('dy1', n * p.dy1_bytes_per_video)
            dy0 = self._workspace('dy0', n * p.dy0_bytes_per_video)
            col_per = max(0 if (self.direct_dgrad1 and self.direct_dgrad0) else p.col0_bytes_per_video, p.col2_bytes_per_video,
                          0 if self.direct_dgrad1 else p.col1_bytes_per_video)
            col = self._workspace('col', n * col_per)
            st = _lib.stream()
            plan = ctypes.byref(p)
            _lib.check(lib.vd_tc_bwd_emb(_lib.ptr(g_emb[s:e]), _lib.ptr(c2[s:e]), _lib.ptr(dy2), plan, n, st), 'vd_tc_bwd_emb')
            _lib.check(lib.vd_tc_bwd_gemm(2, _lib.ptr(dy2), _lib.ptr(self.wt2), _lib.ptr(col), plan, n, st), 'vd_tc_bwd_gemm(2)')
            if self.direct_dgrad1 and self.direct_dgrad0:
                # conv 1 and conv 0: column-free dgrads (shifted-window GEMMs over padded planar dY tensors); the
                # routing of conv 0 is fused into conv 1's epilogue, conv 0's epilogue writes d video
                dyp1 = self._workspace('dyp1', n * self.dyp1_bytes_per_video, zero=True)     # halo cells stay zero
                dyp0 = self._workspace('dyp0', n * self.dyp0_bytes_per_video, zero=True)
                _lib.check(lib.vd_tc_bwd_col2im_ex(2, _lib.ptr(col), _lib.ptr(c1[s:e]), _lib.ptr(dyp1), plan, n, 1, st), 'vd_tc_bwd_col2im_ex(2)')
                _lib.check(lib.vd_tc_dgrad1_ex(_lib.ptr(dyp1), _lib.ptr(self.dg1_w[0]), _lib.ptr(self.dg1_w[1]), _lib.ptr(c0[s:e]),
                                               _lib.ptr(dyp0), plan, n, 1, st), 'vd_tc_dgrad1_ex')
                _lib.check(lib.vd_tc_dgrad0(_lib.ptr(dyp0), _lib.ptr(self.dg0_w), _lib.ptr(dvideo[s:e]), plan, n, 0, st), 'vd_tc_dgrad0')
                continue
            if self.direct_dgrad1:
                dyp1 = self._workspace('dyp1', n * self.dyp1_bytes_per_video, zero=True)
                _lib.check(lib.vd_tc_bwd_col2im_ex(2, _lib.ptr(col), _lib.ptr(c1[s:e]), _lib.ptr(dyp1), plan, n, 1, st), 'vd_tc_bwd_col2im_ex(2)')
                _lib.check(lib.vd_tc_dgrad1(_lib.ptr(dyp1), _lib.ptr(self.dg1_w[0]), _lib.ptr(self.dg1_w[1]), _lib.ptr(c0[s:e]),
                                            _lib.ptr(dy0), plan, n, st), 'vd_tc_dgrad1')
            else:
                _lib.check(lib.vd_tc_bwd_col2im(2, _lib.ptr(col), _lib.ptr(c1[s:e]), _lib.ptr(dy1), plan, n, st), 'vd_tc_bwd_col2im(2)')
                _lib.check(lib.vd_tc_bwd_gemm(1, _lib.ptr(dy1), _lib.ptr(self.wt1), _lib.ptr(col), plan, n, st), 'vd_tc_bwd_gemm(1)')
                _lib.check(lib.vd_tc_bwd_col2im(1, _lib.ptr(col), _lib.ptr(c0[s:e]), _lib.ptr(dy0), plan, n, st), 'vd_tc_bwd_col2im(1)')
            _lib.check(lib.vd_tc_bwd_gemm(0, _lib.ptr(dy0), _lib.ptr(self.wt0), _lib.ptr(col), plan, n, st), 'vd_tc_bwd_gemm(0)')
            _lib.check(lib.vd_tc_bwd_col2im(0, _lib.ptr(col), None, _lib.ptr(dvideo[s:e]), plan, n, st), 'vd_tc_bwd_col2im(0)')
        return dvideo

    def _embed_backward_split(self, g_emb, codes):
        """Backward of the f16x3 modes: route-scatter (fp32) + tensor-core dgrad of each conv evaluated as split-bf16,
        gx = dgrad(gh, wh) + dgrad(gh, wl) + dgrad(gl, wh) with g = gh + gl, w = wh + wl (bf16 parts, fp32 accumulate):
        ~16 significand bits per operand, no range limits on the gradient.  The gradient parts are made inside the packers,
        the weight images of wh / wl are packed once per network, and the column-free dgrads of conv 1 / conv 0 accumulate
        their three passes into one fp32 tensor (conv 0 straight into the (B,T,3,H,W) result).  Returns d video."""
        from . import ops
        from .tc_trio import TcTrio
        if self._trio is None:
            self._trio = TcTrio(self.T, self.H, self.W, self.device)
        trio, p = self._trio, self.plan
        if self._split_bwd_w is None:
            ws = self._fp32_w
            wh = [w.to(torch.bfloat16).float() for w in ws]
            wl = [w - h for w, h in zip(ws, wh)]
            self._split_bwd_w = [(trio.pack_dgrad_weights(l, wh[l]), trio.pack_dgrad_weights(l, wl[l])) for l in range(3)]
        g_emb = g_emb.contiguous().float()
        B = g_emb.shape[0]
        out = torch.empty(B, self.T, 3, self.H, self.W, dtype=torch.float32, device=self.device)
        ext_conv = [(p.T1, p.H1, p.W1), (p.T2, p.H2, p.W2), (p.T3, p.H3, p.W3)]        # conv outputs (pre-pool)
        cout = (64, 128, 128)
        pool = [(1, 2, 2), (2, 2, 2), (2, 2, 2)]
        for s in range(0, B, self.bwd_chunk):
            e = min(B, s + self.bwd_chunk)
            g = g_emb[s:e].view(e - s, 128, p.T3p, p.H3p, p.W3p)
            for layer in (2, 1, 0):
                gy = ops.route_scatter_raw(g, codes[layer][s:e], (e - s, cout[layer]) + tuple(ext_conv[layer]), pool[layer])
                pk_h, pk_l = self._split_bwd_w[layer]
                if layer == 0:
                    # conv 0: the three passes accumulate in the kernel's epilogue, straight into the (B,T,3,H,W) result
                    g = trio.dgrad(0, gy, None, part=0, wpack=pk_h, out=out[s:e], ncdhw=False)
                    trio.dgrad(0, gy, None, part=0, wpack=pk_l, out=g, accumulate=True, ncdhw=False, packed=True)   # gh is still packed
                    trio.dgrad(0, gy, None, part=1, wpack=pk_h, out=g, accumulate=True, ncdhw=False)
                else:
                    # conv 1 / conv 2: separate outputs added by elementwise launches (a read-modify-write in conv 1's epilogue
                    # is a chain of dependent loads per row: 0.52 ms per launch against 0.17 ms for the plain store, measured
                    # again with the row-assembling epilogue)
                    g = trio.dgrad(layer, gy, None, part=0, wpack=pk_h)
                    g += trio.dgrad(layer, gy, None, part=0, wpack=pk_l, packed=True)
                    g += trio.dgrad(layer, gy, None, part=1, wpack=pk_h)
        return out

    def embed_autograd(self, video):
        """Differentiable embed (gradient flows to ``video`` only; the net is frozen as in DM)."""
        return _TcEmbed.apply(video, self)

    def _buffers(self, n, n_total=None):
        """a1 for one chunk of ``n`` videos, a2 for ``n_total`` (>= n) videos: conv 2 runs ONCE over all
        chunks of an embed call (its tiles are 4 videos wide, so small launches waste most of a wave)."""
        n_total = n if n_total is None else n_total
        n4 = (n_total + 3) // 4 * 4
        if self._a1 is None or self._a1.numel() < n * self.a1_per:
            self._a1 = torch.zeros(n * self.a1_per, dtype=torch.uint8, device=self.device)
        if self._a2 is None or self._a2.numel() < n4 * self.a2_per:
            self._a2 = torch.zeros(n4 * self.a2_per, dtype=torch.uint8, device=self.device)
        return self._a1, self._a2

    def set_normalization(self, mean, std):
        """Dataset normalisation applied by the uint8 packer: v = (u/255 - mean[c]) / std[c] (utils.py:214-230)."""
        self._norm = ((ctypes.c_float * 3)(*[float(m) for m in mean]), (ctypes.c_float * 3)(*[float(s) for s in std]))

    def pack_video(self, video, index=None, out=None, hi_only=False):
        """fp32 (or uint8 frames, see set_normalization) (Bsrc,T,3,H,W) -> X0 for B = len(index) (or Bsrc) items.
        hi_only (split mode): the hi-only operand X0h of the two-product mode."""
        assert video.dtype in (torch.float32, torch.uint8) and video.dim() == 5 and tuple(video.shape[1:]) == (self.T, 3, self.H, self.W)
        assert not hi_only or self.split
        B = int(index.numel()) if index is not None else int(video.shape[0])
        nbytes = B * (self.x0h_per if hi_only else self.x0_per)
        lib = _lib.lib()
        if out is None:
            if self._x0 is None or self._x0.numel() < nbytes:
                self._x0 = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            out = self._x0
        if video.dtype == torch.uint8:
            if getattr(self, '_norm', None) is None:
                raise RuntimeError('uint8 videos need TcConvNet3D.set_normalization(mean, std) first')
            fn = (lib.vd_tc_x3_pack_video_hi_u8 if hi_only else lib.vd_tc_x3_pack_video_u8) if self.split else lib.vd_tc_pack_video_u8
            _lib.check(fn(_lib.ptr(video), _lib.ptr(index), _lib.ptr(out), ctypes.byref(self.plan), B,
                          self._norm[0], self._norm[1], _lib.stream()), 'vd_tc_pack_video_u8')
            return out
        fn = (lib.vd_tc_x3_pack_video_hi if hi_only else lib.vd_tc_x3_pack_video) if self.split else lib.vd_tc_pack_video
        _lib.check(fn(_lib.ptr(video), _lib.ptr(index), _lib.ptr(out), ctypes.byref(self.plan), B, _lib.stream()), 'vd_tc_pack_video')
        return out

    # ---------------------------------------------------------------- layers
    def conv_layer(self, layer, src, wimg, bias, out, B, code=None, item_index=None, raw=False, code_first=0, products=3):
        ev = None
        if self.timing is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self._conv_layer(layer, src, wimg, bias, out, B, code, item_index, raw, code_first, products)
        if ev is not None:
            ev[1].record()
            self.timing.append((layer, int(B), ev[0], ev[1], int(products) if self.split else 1))

    def _conv_layer(self, layer, src, wimg, bias, out, B, code=None, item_index=None, raw=False, code_first=0, products=3):
        if self.split:
            assert not raw, 'the split-fp16 path has fused epilogues only'
            assert products == 3 or code is None, 'the two-product mode is for frozen videos (no routing codes)'
            _lib.check(_lib.lib().vd_tc_x3_conv_layer_ex(layer, _lib.ptr(src), _lib.ptr(wimg), _lib.ptr(bias), _lib.ptr(out),
                                                         _lib.ptr(code), int(code_first), ctypes.byref(self.plan), _lib.ptr(item_index),
                                                         int(B), int(products), _lib.stream()), f'vd_tc_x3_conv_layer_ex({layer})')
            return
        _lib.check(_lib.lib().vd_tc_conv_layer(layer, _lib.ptr(src), _lib.ptr(wimg), _lib.ptr(bias), _lib.ptr(out),
                                               _lib.ptr(code), int(code_first), ctypes.byref(self.plan), _lib.ptr(item_index),
                                               int(B), int(raw), _lib.stream()), f'vd_tc_conv_layer({layer})')

    def embed_packed(self, x0, B, item_index=None, out=None, codes=None):
        """conv0..2 on an already packed X0 (optionally a resident set addressed by item_index)."""
        a1, a2 = self._buffers(B)
        if out is None:
            out = torch.empty(B, self.embed_dim, dtype=torch.float32, device=self.device)
        c0, c1, c2 = codes if codes is not None else (None, None, None)
        self.conv_layer(0, x0, self.w0, self.b0, a1, B, code=c0, item_index=item_index)
        self.conv_layer(1, a1, self.w1, self.b1, a2, B, code=c1)
        self.conv_layer(2, a2, self.w2, self.b2, out, B, code=c2)
        return out

    def _embed_chunks(self, B, out, codes, front, code_first=0, products=3):
        """conv 0 + conv 1 per chunk of max_batch videos (``front(s, e)`` -> (x0, item_index)), every
        chunk writing its slice of one A2 buffer; then ONE conv-2 launch over all B videos.  ``codes``
        (optional) hold the routing of items code_first..B-1 only."""
        p = self.plan
        step = max(4, self.max_batch // 4 * 4)              # conv-2 tiles hold 4 consecutive videos
        n_chunks = (B + step - 1) // step
        step = min(step, ((B + n_chunks - 1) // n_chunks + 3) // 4 * 4)      # equal chunks: no small, wave-inefficient tail
        a1, a2 = self._buffers(min(B, step), B)
        c0, c1, c2 = codes if codes is not None else (None, None, None)
        for s in range(0, B, step):
            e = min(B, s + step)
            x0, idx = front(s, e)
            cc0 = cc1 = None
            first = 0
            if c0 is not None and e > code_first:
                first = max(0, code_first - s)              # chunk-local index of the first item with codes
                cc0, cc1 = c0[s + first - code_first:], c1[s + first - code_first:]
            self.conv_layer(0, x0, self.w0, self.b0, a1, e - s, code=cc0, item_index=idx, code_first=first, products=products)
            self.conv_layer(1, a1, self.w1, self.b1, a2[s * self.a2_per:], e - s, code=cc1, code_first=first, products=products)
        self.conv_layer(2, a2, self.w2, self.b2, out, B, code=c2, code_first=code_first, products=products)
        return out

    def alloc_codes(self, B):
        p = self.plan
        u8 = dict(dtype=torch.uint8, device=self.device)
        return (torch.empty(B, 64, p.T1p, p.H1p, p.W1p, **u8), torch.empty(B, 128, p.T2p, p.H2p, p.W2p, **u8),
                torch.empty(B, 128, p.T3p, p.H3p, p.W3p, **u8))

    def pack_dataset(self, videos, chunk=256, extra_slots=0):
        """One-time conversion of a resident fp32 real set (N,T,3,H,W) into the packed conv-0 operand
        (bf16, 'kw-expanded'); afterwards ``embed_resident`` reads it in place through item_index.
        ``extra_slots`` spare video slots at the tail receive the synthetic videos of ``embed_joint``."""
        N = int(videos.shape[0])
        hi_only = self.real_products == 2        # frozen set in the two-product mode: hi part only, half the bytes, no spare slots
        per = self.x0h_per if hi_only else self.x0_per
        x0 = torch.empty((N + (0 if hi_only else extra_slots)) * per, dtype=torch.uint8, device=self.device)
        for s in range(0, N, chunk):
            e = min(N, s + chunk)
            self.pack_video(videos[s:e], out=x0[s * per:e * per], hi_only=hi_only)
        return x0

    def embed_resident(self, x0_all, index):
        """embed of the videos ``index`` (device int64) of a pre-packed resident set."""
        B = int(index.numel())
        out = torch.empty(B, self.embed_dim, dtype=torch.float32, device=self.device)
        index = index.contiguous()
        return self._embed_chunks(B, out, None, lambda s, e: (x0_all, index[s:e]), products=self.real_products)

    def embed_joint(self, x0_all, index_real, video_syn, tail_slot):
        """Frozen real videos (resident, addressed by ``index_real``) and differentiable synthetic videos in ONE
        pass of the three conv kernels: the synthetic videos are packed into the spare slots ``tail_slot...`` of
        the resident operand and only they record routing codes.  Returns (emb_real, emb_syn, codes)."""
        n_real, n_syn = int(index_real.numel()), int(video_syn.shape[0])
        if self.real_products == 2:
            # two launches per layer: the frozen real videos in the two-product mode on the hi-only resident operand, the
            # differentiable synthetic videos (1.5 % of the work at ipc = 1) with all three products and routing codes
            emb_real = self.embed_resident(x0_all, index_real.reshape(-1))
            emb_syn, codes = self.embed(video_syn.contiguous(), want_codes=True)
            return emb_real, emb_syn, codes
        per = self.x0_per
        assert x0_all.numel() >= (tail_slot + n_syn) * per, 'resident operand has no spare slots for the synthetic videos'
        self.pack_video(video_syn.contiguous(), out=x0_all[tail_slot * per:(tail_slot + n_syn) * per])
        index = torch.cat([index_real.reshape(-1), torch.arange(tail_slot, tail_slot + n_syn, device=self.device)])
        codes = self.alloc_codes(n_syn)
        out = torch.empty(n_real + n_syn, self.embed_dim, dtype=torch.float32, device=self.device)
        self._embed_chunks(n_real + n_syn, out, codes, lambda s, e: (x0_all, index[s:e]), code_first=n_real)
        return out[:n_real], out[n_real:].clone(), codes

    def embed_joint_autograd(self, x0_all, index_real, video_syn, tail_slot):
        """(emb_real [no grad], emb_syn [differentiable w.r.t. video_syn]) of one joint pass."""
        emb_syn, emb_real = _TcEmbedJoint.apply(video_syn, self, x0_all, index_real, tail_slot)
        return emb_real, emb_syn

    def embed(self, video, index=None, want_codes=False, frozen=False):
        """ConvNet3D.embed on fp32 videos (B,T,3,H,W) -> (B, embed_dim) fp32, in chunks of max_batch.
        frozen=True (no codes): these are real videos whose embeddings are only averaged — two-product mode if enabled."""
        B = int(index.numel()) if index is not None else int(video.shape[0])
        out = torch.empty(B, self.embed_dim, dtype=torch.float32, device=self.device)
        codes = self.alloc_codes(B) if want_codes else None
        hi_only = frozen and not want_codes and self.real_products == 2

        def front(s, e):
            if index is not None:
                return self.pack_video(video, index[s:e], hi_only=hi_only), None
            return self.pack_video(video[s:e], hi_only=hi_only), None
        self._embed_chunks(B, out, codes, front, products=2 if hi_only else 3)
        return (out, codes) if want_codes else out


class _TcEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, video, net):
        emb, codes = net.embed(video.contiguous(), want_codes=True)
        ctx.net, ctx.codes = net, codes
        return emb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_emb):
        return ctx.net.embed_backward(g_emb, ctx.codes), None


class _TcEmbedJoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, video_syn, net, x0_all, index_real, tail_slot):
        emb_real, emb_syn, codes = net.embed_joint(x0_all, index_real, video_syn, tail_slot)
        ctx.net, ctx.codes = net, codes
        ctx.mark_non_differentiable(emb_real)
        return emb_syn, emb_real

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_syn, g_real):
        return ctx.net.embed_backward(g_syn, ctx.codes), None, None, None, None
