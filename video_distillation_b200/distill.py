"""The distillation inner loops re-expressed for one-process-per-GPU execution.

* ``DeviceDataset``   — the real set resident in HBM (class-sharded across ranks); ``get_images``
  keeps the reference's numpy sampling (distill_s2d_ms.py:81-87) but gathers on the device.
* ``DMS2DTrainer``    — DM + static/dynamic memory (distill_s2d_ms.py:393-438).
* ``DMBaselineTrainer`` — DM on leaf synthetic videos (distill_baseline.py:334-356).
* ``MTTS2DTrainer``   — MTT unrolled student over flat parameters (distill_s2d_ms.py:197-300).

The 50..400 per-class Python iterations of the reference become a handful of batched launches:
all real videos of the rank's classes are embedded in one pass (tensor cores, bf16 operands,
fp32 accumulate, or the exact fp32 path with precision='fp32'), reduced to class means, and the
loss + d loss/d embedding come out of one kernel.  Multi-GPU: the synthetic branch and the memories
(with their momentum buffers) are sharded by class ``c % world == rank``; the sampled real videos of
every class are spread over the ranks (``DeviceDataset(shard='video')``) and completed by an all-reduce
of the (C, D) partial embedding sums; every rank replays the full RNG streams and slices its part, so
sampling is bit-identical to a single-GPU run.  Per iteration two small collectives cross NVLink:
an all-gather of the real embeddings (C*n*D floats, reduced in single-GPU order: class means bitwise
independent of the world size; `exact_means=False` all-reduces (C, D) partial sums instead) and
[hallucinator grad | loss] (328 floats).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .networks import ConvNet3D
from .reparam_module import ReparamModule
from .tc import TcConvNet3D, tc_supported
from .utils import Conv3DNet, get_network


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def owned_classes(num_classes, rank, world):
    """Class shard of a rank: c % world == rank.  50 classes / 8 ranks -> 7,7,6,6,6,6,6,6."""
    return [c for c in range(num_classes) if c % world == rank]


def allreduce_sum_(tensors):
    """SUM all-reduce of a list of tensors, in place; no-op for a single rank.  Tensors below 64 Ki elements travel as ONE
    flat message, larger contiguous ones are reduced in place one by one (no staging copy).  The DM path calls this once per
    iteration with [hallucinator grads, loss] (a single 328-float message: the memory gradients are class-local)."""
    rank, world = _world()
    if world == 1:
        return tensors
    small = [t for t in tensors if t.numel() < (1 << 16) or not t.is_contiguous()]
    for t in tensors:
        if t.numel() >= (1 << 16) and t.is_contiguous():
            dist.all_reduce(t, op=dist.ReduceOp.SUM)            # big (dynamic-memory grad): in place, no staging copy
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        ofs = 0
        for t in small:
            t.copy_(flat[ofs:ofs + t.numel()].view_as(t))
            ofs += t.numel()
    return tensors


class _SumAcrossRanks(torch.autograd.Function):
    """y = sum over ranks of x_r (all-reduce).  y feeds a computation that every rank replicates, so the
    cotangent that arrives on a rank is already the complete one: backward is the identity."""
    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        return g


class _ReplicatedInput(torch.autograd.Function):
    """Identity on a replicated tensor (the student parameters theta_k) that enters rank-local work (the rank's
    shard of the step batch).  Every rank produces only its shard's part of d/d theta_k, so the backward sums the
    cotangents over ranks (all-reduce): the complete cotangent continues into the earlier unroll steps."""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def sharded_inner_grad(theta, shard_loss_fn, world):
    """grad_theta of the FULL step loss when each rank evaluates only its shard:
    ``shard_loss_fn(theta_local)`` must return sum_{i in shard} loss_i / B_total (0-dim, built on theta_local).
    Returns the all-reduced gradient, differentiable to second order across ranks (MTT row)."""
    if world == 1:
        return torch.autograd.grad(shard_loss_fn(theta), theta, create_graph=True)[0]
    theta_local = _ReplicatedInput.apply(theta)
    g_local = torch.autograd.grad(shard_loss_fn(theta_local), theta_local, create_graph=True)[0]
    return _SumAcrossRanks.apply(g_local)


class DeviceDataset:
    """Real videos of this rank's classes, resident on the device.

    ``videos`` (N, T, 3, H, W) fp32 and ``labels`` (N,) describe the FULL dataset exactly like the
    reference's preloaded TensorDataset (distill_s2d_ms.py:28-38); only the rows whose class is
    owned by this rank (``c % world == rank``) are uploaded.  ``indices_class`` is built as at
    distill_s2d_ms.py:73-79, so ``np.random.permutation(indices_class[c])[:n]`` is unchanged.
    """

    def __init__(self, videos, labels, num_classes, device, rank=0, world=1, norm=None, shard='class'):
        # ``videos`` may be the decoded uint8 frames with ``norm = (mean, std)``: the tensor-core path then normalises inside
        # its packer ((u/255 - mean)/std, bit-identical operands) and the resident set costs 1 byte per element
        self.norm = norm
        if videos.dtype == torch.uint8 and norm is None:
            raise ValueError('uint8 videos need norm=(mean, std)')
        labels = [int(v) for v in labels]
        self._index(labels, num_classes, rank, world, shard)
        keep = self.keep
        self.local_of_global = np.full(len(labels), -1, dtype=np.int64)
        self.local_of_global[np.asarray(keep, dtype=np.int64)] = np.arange(len(keep))
        self.device = torch.device(device)
        idx = torch.as_tensor(keep, dtype=torch.long)
        self.videos = videos[idx].to(self.device, non_blocking=True).contiguous()
        self.shape = tuple(videos.shape[1:])
        self.x0 = None

    def _index(self, labels, num_classes, rank, world, shard):
        """indices_class as at distill_s2d_ms.py:73-79 and the rows this rank holds:
        shard='class': the videos of the classes c % world == rank (the rank embeds whole classes);
        shard='video': the videos whose POSITION inside their class is p % world == rank — every class is spread evenly over
                       the ranks, so a draw of batch_real videos per class gives every rank ~batch_real / world of EVERY class
                       (hypergeometric, +-1) and the real-video work balances to ~1 % whatever C / world is; the per-class
                       embedding sums are completed by an all-reduce of (C, D) partial sums."""
        if shard not in ('class', 'video'):
            raise ValueError("shard must be 'class' or 'video'")
        self.shard = shard
        self.num_classes = num_classes
        self.indices_class = [[] for _ in range(num_classes)]
        pos = []
        for i, lab in enumerate(labels):
            pos.append(len(self.indices_class[lab]))
            self.indices_class[lab].append(i)
        self.rank, self.world = rank, world
        self.pos_in_class = np.asarray(pos, dtype=np.int64)
        self.owned = owned_classes(num_classes, rank, world)
        if shard == 'class':
            self.keep = [i for i, lab in enumerate(labels) if lab % world == rank]
        else:
            self.keep = [i for i in range(len(labels)) if pos[i] % world == rank]

    @classmethod
    def from_device_shard(cls, shard_videos, labels, num_classes, device, rank=0, world=1, shard='class'):
        """Same object from an already device-resident shard: ``shard_videos`` holds exactly the rows this rank keeps
        (``_index``), in dataset order (synthetic benchmarks generate them on the device)."""
        self = cls.__new__(cls)
        labels = [int(v) for v in labels]
        self._index(labels, num_classes, rank, world, shard)
        keep = self.keep
        assert shard_videos.shape[0] == len(keep), 'shard does not match the owned rows'
        self.local_of_global = np.full(len(labels), -1, dtype=np.int64)
        self.local_of_global[np.asarray(keep, dtype=np.int64)] = np.arange(len(keep))
        self.device = torch.device(device)
        self.videos = shard_videos.contiguous()
        self.shape = tuple(shard_videos.shape[1:])
        self.x0 = None
        self.norm = None
        return self

    def prepack(self, tc_net, free_fp32=False, extra_slots=0):
        """Convert the resident set once into the tensor-core path's packed bf16 conv-0 operand
        (device-resident real-data pipeline); per-iteration packing disappears.
        ``extra_slots`` spare slots let the trainer embed its synthetic videos in the same launches."""
        if self.videos.dtype == torch.uint8:
            tc_net.set_normalization(*self.norm)
        self.x0 = tc_net.pack_dataset(self.videos, extra_slots=extra_slots)
        self.x0_tail = int(self.videos.shape[0])
        self.x0_extra = int(extra_slots)
        if free_fp32:
            self.videos = None
        return self

    def sample_all_classes(self, n):
        """The reference's per-class draws for ALL classes in order (every rank replays the whole
        numpy stream); returns the global indices (C, n) — bit-exact with get_images."""
        return np.stack([np.random.permutation(self.indices_class[c])[:n] for c in range(self.num_classes)])

    def local_sample(self, real_idx):
        """shard='video': of the (C, n) sampled global indices of ALL classes, the ones this rank holds — as rows of
        ``self.videos`` in class-major order (sampled order inside a class) plus the (C+1,) int32 segment offsets."""
        loc_all = self.local_of_global[np.asarray(real_idx)]
        mask = loc_all >= 0
        offsets = np.zeros(loc_all.shape[0] + 1, dtype=np.int32)
        np.cumsum(mask.sum(1), out=offsets[1:])
        return self._to_device(np.ascontiguousarray(loc_all[mask])), self._to_device(offsets)

    def gather_plan(self, real_idx):
        """shard='video': how the ranks' local shares of a (C, n) draw reassemble into sample order.  Every rank can compute
        this for ALL ranks (the draw and the position rule are replicated).  Returns (n_max, index) where n_max is the largest
        local share and index (device int64, C*n) addresses the (world * n_max, D) all-gathered embedding rows such that
        gathered[index].view(C, n, D) is the embedding of real_idx[c, j] — the order a single GPU computes them in."""
        real_idx = np.asarray(real_idx)
        C, n = real_idx.shape
        owner = self.pos_in_class[real_idx] % self.world                   # (C, n) rank holding each sampled video
        flat_owner = owner.reshape(-1)
        counts = np.bincount(flat_owner, minlength=self.world)
        n_max = int(counts.max())
        # a rank's local rows are in class-major / sample order = the order of appearance in the flattened draw
        rank_pos = np.zeros(C * n, dtype=np.int64)
        for r in range(self.world):
            m = flat_owner == r
            rank_pos[m] = np.arange(int(m.sum()))
        return n_max, self._to_device(np.ascontiguousarray(flat_owner.astype(np.int64) * n_max + rank_pos))

    def local_index(self, global_idx):
        """global video indices of owned classes -> rows of ``self.videos`` (device int64)."""
        loc = self.local_of_global[np.asarray(global_idx).reshape(-1)]
        assert (loc >= 0).all(), 'requested a video this rank does not hold'
        return self._to_device(np.ascontiguousarray(loc))

    def _to_device(self, arr):
        t = torch.from_numpy(arr)
        if self.device.type != 'cuda':
            return t.to(self.device)
        # ring of reusable pinned staging buffers: a fresh pin_memory() per call is a cudaHostAlloc, which
        # synchronises the device (0.5 ms idle per iteration in the timeline); 8 slots keep in-flight copies apart
        ring = getattr(self, '_pin_ring', None)
        nbytes = t.numel() * t.element_size()
        if ring is None or ring[0].numel() < nbytes:
            ring = self._pin_ring = [torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory() for _ in range(8)]
            self._pin_done = [None] * len(ring)
            self._pin_next = 0
        slot = self._pin_next % len(ring)
        self._pin_next += 1
        if self._pin_done[slot] is not None:
            self._pin_done[slot].synchronize()            # the copy that last read this slot (8 calls ago) has finished
        buf = ring[slot][:nbytes].view(t.dtype)
        buf.copy_(t)
        out = buf.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pin_done[slot] = ev
        return out

    def get_images(self, c, n):
        """Reference-compatible accessor (one class)."""
        idx = np.random.permutation(self.indices_class[c])[:n]
        v = self.videos[self.local_index(idx)]
        if v.dtype == torch.uint8:                                         # the normalised floats the reference would hold
            mean = torch.tensor(self.norm[0], device=v.device).view(1, 1, 3, 1, 1)
            std = torch.tensor(self.norm[1], device=v.device).view(1, 1, 3, 1, 1)
            # tensor / tensor: IEEE division like the host transform (a Python-scalar divisor becomes a reciprocal multiply on CUDA)
            v = ((v.float() / torch.tensor(255.0, device=v.device)) - mean) / std
        return v


def frozen_convnet3d(channel, num_classes, im_size, frames, device, seed=None, init_on_device=False):
    """A fresh frozen random ConvNet3D as at distill_s2d_ms.py:393-396.  ``seed`` (tests, multi-GPU:
    every rank must build the same net) replaces the wall-clock reseed of get_network.
    ``init_on_device`` draws the same default init (kaiming-uniform / uniform bias) with the DEVICE
    generator instead of initialising 3.65 M parameters on the host and copying them every
    iteration (same distribution, different stream than the reference's CPU init)."""
    if init_on_device:
        if seed is not None:
            torch.cuda.manual_seed(int(seed))
        with torch.device(device):
            net = ConvNet3D(channel, num_classes, 128, 3, 'relu', 'none', 'maxpooling', frames, im_size)
    elif seed is None:
        net = get_network('ConvNet3D', channel, num_classes, im_size, frames=frames)
    else:
        torch.random.manual_seed(int(seed))
        net = ConvNet3D(channel, num_classes, 128, 3, 'relu', 'none', 'maxpooling', frames, im_size).to(device)
    net.train()
    for p in net.parameters():
        p.requires_grad = False
    return net


class FrozenNetPool:
    """The "fresh random frozen ConvNet3D per iteration" (distill_s2d_ms.py:393-396) without rebuilding
    a module every step: one persistent ConvNet3D whose parameters are views of a flat device buffer;
    ``fresh(seed)`` redraws the default init (weights and biases ~ U(+-1/sqrt(fan_in)), which is what
    kaiming_uniform(a=sqrt 5) + the bias rule of nn.Conv3d reduce to) with two launches on the device
    generator.  Same distribution as get_network, different random stream (device vs host)."""

    def __init__(self, channel, num_classes, im_size, frames, device):
        with torch.device(device):
            self.net = ConvNet3D(channel, num_classes, 128, 3, 'relu', 'none', 'maxpooling', frames, im_size)
        params = list(self.net.parameters())
        self.flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=device)
        self.scale = torch.empty_like(self.flat)
        ofs = 0
        convs = [m for m in self.net.modules() if isinstance(m, torch.nn.Conv3d)]
        for m in convs:
            fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3] * m.weight.shape[4]
            for p in (m.weight, m.bias):
                n = p.numel()
                self.scale[ofs:ofs + n] = 1.0 / (fan_in ** 0.5)
                p.data = self.flat[ofs:ofs + n].view_as(p)
                p.requires_grad = False
                ofs += n
        assert ofs == self.flat.numel()
        self.net.train()

    def fresh(self, seed=None):
        if seed is not None:
            torch.cuda.manual_seed(int(seed))
        self.flat.uniform_(-1.0, 1.0)
        self.flat.mul_(self.scale)
        return self.net


class _RealEmbedder:
    """Embeds real videos (forward only, frozen net) on tensor cores or on the exact fp32 path."""

    def __init__(self, frames, im_size, device, precision, max_batch):
        self.precision = precision
        self.device = device
        self.max_batch = max_batch
        self.tc = None
        if precision in ('bf16', 'f16x3', 'f16x3r2'):
            # 'f16x3': the fused pipeline on fp16 hi/lo operand pairs (fp32-equivalent embeddings and routing, the parity
            # mode of the fast path); 'f16x3r2': the same, with the frozen real videos in the two-product mode (exact weights,
            # activations rounded once to fp16 — tc.TcConvNet3D); 'bf16': single-pass bf16 operands and activations
            if not tc_supported(frames, im_size[0], im_size[1]):
                raise RuntimeError(f'tensor-core path does not support videos {frames}x{im_size}')
            self.tc = TcConvNet3D(frames, im_size[0], im_size[1], device, max_batch=max_batch, split=precision.startswith('f16x3'),
                                  real_products=2 if precision == 'f16x3r2' else 3)
        elif precision not in ('fp32', 'bf16x3'):
            raise ValueError("precision must be 'f16x3r2', 'f16x3', 'bf16', 'bf16x3' or 'fp32'")

    def load(self, net):
        self.net = net
        if self.tc is not None:
            f = net.features
            self.tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)

    @torch.no_grad()
    def __call__(self, videos, index, x0=None):
        if self.tc is not None:
            if x0 is not None:
                return self.tc.embed_resident(x0, index)
            return self.tc.embed(videos, index=index, frozen=True)
        # 'fp32': exact CUDA-core kernels.  'bf16x3': the same module on the tensor-core conv trio with split-bf16
        # fprop and fp32 activations (embeddings ~1e-5 of fp32) — the caller holds ops.set_conv_backend('tc').
        out = []
        chunk = min(self.max_batch, 64) if self.precision == 'bf16x3' else self.max_batch
        for s in range(0, index.numel(), chunk):
            out.append(self.net.embed(videos[index[s:s + chunk]]))
        return torch.cat(out, 0)


class DMS2DTrainer:
    """State + iteration of DM with static/dynamic memory (distill_s2d_ms.py:89-108, 393-438)."""

    def __init__(self, dataset, *, num_classes, channel=3, im_size=(112, 112), frames=16, vpc=1, spc=2, dpc=2,
                 batch_real=64, lr_dynamic=1e4, lr_hal=1e-2, lr_static=1e-4, train_static=False, precision='bf16',
                 static_syn=None, dynamic_syn=None, hal=None, max_batch=128, device='cuda', init_on_device=False,
                 syn_on_tensor_cores=True):
        self.rank, self.world = _world()
        self.ds = dataset
        self.init_on_device = init_on_device
        self._pool = None
        self._const = None
        # precision='bf16', synthetic branch:
        #   True    fused tensor-core pipeline (bf16 operands AND bf16 activations between layers; throughput mode —
        #           ReLU / pool routing can flip against fp32, ~1e-1 relL2 on the synthetic-video gradient)
        #   'split' tensor-core conv trio with split-bf16 fprop and fp32 activations (routing identical to fp32 up to
        #           ~1e-5, gradients ~1e-2), unfused: a few ms per 50 videos
        #   False   exact fp32 CUDA-core kernels (slow)
        self.syn_on_tensor_cores = syn_on_tensor_cores
        self.C, self.channel, self.im_size, self.frames = num_classes, channel, tuple(im_size), frames
        self.vpc, self.spc, self.dpc, self.batch_real = vpc, spc, dpc, batch_real
        self.lr_dynamic, self.lr_hal, self.lr_static, self.train_static = lr_dynamic, lr_hal, lr_static, train_static
        self.device = torch.device(device)
        H, W = self.im_size
        # distill_s2d_ms.py:89-93 (CPU randn then .to(device), so seeds reproduce the reference)
        if static_syn is None:
            static_syn = torch.randn(size=(num_classes * spc, 3, H, W), dtype=torch.float)
        if dynamic_syn is None:
            dynamic_syn = torch.randn(size=(num_classes, dpc, frames, 1, H, W), dtype=torch.float)
        self.hal = (hal if hal is not None else Conv3DNet()).to(self.device)
        self.owned = owned_classes(num_classes, self.rank, self.world)
        self.owned_t = torch.as_tensor(self.owned, dtype=torch.long, device=self.device)
        # Multi-GPU: the memories (and their momentum buffers) are SHARDED by class — a class's synthetic videos only ever read
        # and update that class's rows (distill_s2d_ms.py:402-412), so a rank keeps the rows of its classes c % world == rank
        # and nothing of the 80..520 MB dynamic memory crosses NVLink during training (`full_memories` gathers for eval / save).
        # Every rank receives the same full-size init (same seeds / broadcast) and slices its rows.
        self.sharded_memories = self.world > 1
        self.exact_means = True          # video-sharded real set: all-gather embeddings (bitwise = 1 GPU) instead of partial sums
        if self.sharded_memories:
            own_cpu = torch.as_tensor(self.owned, dtype=torch.long)
            static_syn = static_syn.detach().cpu().view(num_classes, spc, *static_syn.shape[1:])[own_cpu].reshape(-1, *static_syn.shape[1:])
            dynamic_syn = dynamic_syn.detach().cpu()[own_cpu]
        self.static_syn = static_syn.detach().to(self.device).contiguous().requires_grad_(train_static)
        self.dynamic_syn = dynamic_syn.detach().to(self.device).contiguous().requires_grad_(True)
        self._bufs = {}
        self.embedder = _RealEmbedder(frames, self.im_size, self.device, precision, max_batch)
        self.last = {}
        self.timeline = None             # set to [] to collect (phase name, CUDA event) marks of every step (bench.py --timeline)

    def _mark(self, name):
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((name, ev))

    # ------------------------------------------------------------------ pieces
    def sample_syn_indices(self):
        """distill_s2d_ms.py:402-406 verbatim: two device randint draws of length C*vpc."""
        C, vpc, spc, dev = self.C, self.vpc, self.spc, self.device
        if getattr(self, '_const', None) is None:
            # label / idx are the same every iteration: build them once (the two randint draws are not)
            label = torch.tensor(np.stack([np.ones(vpc) * i for i in range(0, C)]), dtype=torch.long,
                                 requires_grad=False, device=dev).view(-1)
            ran = torch.arange(0, C * vpc).to(dev)
            idx = ran % vpc
            self._const = (label, 2 * idx, spc * label + 2 * idx)
        label, idx2, sbase = self._const
        dynamic_idx = idx2 + torch.randint(2, (C * vpc,), device=dev)
        static_idx = sbase + torch.randint(2, (C * vpc,), device=dev)
        return label, dynamic_idx, static_idx

    def _sgd(self, name, p, grad, lr, momentum=0.95):
        first = name not in self._bufs
        if first:
            self._bufs[name] = torch.empty_like(p)
        ops.sgd_momentum_(p.data, grad.contiguous(), self._bufs[name], lr, momentum, first)

    def full_memories(self):
        """(static_syn (C*spc,3,H,W), dynamic_syn (C,dpc,T,1,H,W)) of ALL classes on every rank (collective when world > 1:
        each rank contributes its class shard).  For evaluation / checkpoints, outside the iteration."""
        if not self.sharded_memories:
            return self.static_syn.detach(), self.dynamic_syn.detach()
        H, W = self.im_size
        st = torch.zeros(self.C, self.spc, 3, H, W, device=self.device)
        dy = torch.zeros(self.C, self.dpc, self.frames, 1, H, W, device=self.device)
        st[self.owned_t] = self.static_syn.detach().view(len(self.owned), self.spc, 3, H, W)
        dy[self.owned_t] = self.dynamic_syn.detach()
        dist.all_reduce(st)
        dist.all_reduce(dy)
        return st.view(self.C * self.spc, 3, H, W), dy

    # ------------------------------------------------------------------ one iteration
    def step(self, net=None, net_seed=None, indices=None, real_idx=None, real_batch=None, real_batch_index=None,
             real_batch_offsets=None):
        """One DM iteration; returns the loss (0-dim device tensor, summed over ALL classes).
        ``real_batch``: optional device tensor holding this rank's sampled real videos already gathered
        (class-major, batch_real per owned class) — the host-streaming mode of bench.py; ``real_batch_index`` (device int64)
        maps sample j to its row of ``real_batch`` when the rows were uploaded in another order (merged host ranges);
        ``real_batch_offsets`` (device int32, C+1): class segments of the batch when the set is video-sharded."""
        if (self.embedder.tc is not None and self.syn_on_tensor_cores == 'split') or self.embedder.precision == 'bf16x3':
            # forward AND backward of net.embed(...) below; precision='bf16x3': every primitive of the trio on hi / lo pairs
            prev = ops.set_conv_backend('tc_x3' if self.embedder.precision == 'bf16x3' else 'tc')
            try:
                return self._step(net, net_seed, indices, real_idx, real_batch, real_batch_index, real_batch_offsets)
            finally:
                ops.set_conv_backend(prev)
        return self._step(net, net_seed, indices, real_idx, real_batch, real_batch_index, real_batch_offsets)

    def _step(self, net=None, net_seed=None, indices=None, real_idx=None, real_batch=None, real_batch_index=None,
              real_batch_offsets=None):
        C, vpc = self.C, self.vpc
        self._mark('begin')
        if net is None:
            if self.init_on_device:
                if self._pool is None:
                    self._pool = FrozenNetPool(self.channel, C, self.im_size, self.frames, self.device)
                net = self._pool.fresh(net_seed)
            else:
                net = frozen_convnet3d(self.channel, C, self.im_size, self.frames, self.device, seed=net_seed)
        self.embedder.load(net)
        self._mark('fresh net + weight images')
        label, dynamic_idx, static_idx = indices if indices is not None else self.sample_syn_indices()
        if real_idx is None:
            real_idx = self.ds.sample_all_classes(self.batch_real)          # (C, batch_real) global, host
        # ---- this rank's classes
        own = self.owned
        n_own = len(own)
        sel = (self.owned_t[:, None] * vpc + torch.arange(vpc, device=self.device)[None, :]).reshape(-1)
        if self.sharded_memories:
            # rows of the class shard: class c = own[k] lives at local position k (static rows spc*k .. spc*k + spc - 1)
            lab_loc = torch.arange(n_own, device=self.device).repeat_interleave(vpc)
            image_syn = self.hal.compose(self.static_syn, self.dynamic_syn, static_idx[sel] - self.spc * (label[sel] - lab_loc),
                                         lab_loc, dynamic_idx[sel], unique_rows=indices is None)
        else:
            # (label, dynamic_idx) pairs of distill_s2d_ms.py:405 are distinct by construction; caller-supplied indices may not be
            image_syn = self.hal.compose(self.static_syn, self.dynamic_syn, static_idx[sel], label[sel], dynamic_idx[sel],
                                         unique_rows=indices is None)
        self._mark('composer forward')
        tc = self.embedder.tc
        fused_syn = tc is not None and self.syn_on_tensor_cores is True
        joint = (fused_syn and real_batch is None and self.ds.x0 is not None
                 and (tc.real_products == 2 or getattr(self.ds, 'x0_extra', 0) >= image_syn.shape[0]))
        video_sharded = self.world > 1 and getattr(self.ds, 'shard', 'class') == 'video'
        offsets = real_batch_offsets
        if real_batch is not None:
            assert not video_sharded or offsets is not None, 'a streamed batch of a video-sharded set needs its class offsets'
        elif video_sharded:
            ridx, offsets = self.ds.local_sample(real_idx)                  # this rank's share of EVERY class's draw
        else:
            ridx = self.ds.local_index(real_idx[own])                       # (n_own*batch_real,)
        if joint:
            # real + synthetic videos in the same three conv launches (codes only for the synthetic tail)
            emb_real, emb_syn = tc.embed_joint_autograd(self.ds.x0, ridx, image_syn, self.ds.x0_tail)
        elif real_batch is None:
            emb_real = self.embedder(self.ds.videos, ridx, x0=self.ds.x0)   # (n_own*batch_real, D)
        else:
            ridx = real_batch_index if real_batch_index is not None else torch.arange(real_batch.shape[0], device=self.device)
            emb_real = self.embedder(real_batch, ridx)
        D = emb_real.shape[1]
        self._mark('embed (conv 0/1/2)')
        if video_sharded and self.exact_means:
            # all-gather the ranks' embedding rows (C*n*D floats: 26 MB at the bench shape, ~0.1 ms over NVLink), put them back
            # in sample order and reduce them with the SAME kernel in the SAME order as a single GPU: the class means — and so
            # the loss, routing and memory gradients — are bitwise independent of the world size
            n_max, gidx = self.ds.gather_plan(real_idx)
            pad = torch.zeros(n_max, D, device=self.device) if emb_real.shape[0] < n_max else None
            mine = emb_real if pad is None else torch.cat([emb_real, pad[:n_max - emb_real.shape[0]]])
            gathered = torch.empty(self.world * n_max, D, device=self.device)
            dist.all_gather_into_tensor(gathered, mine.contiguous())
            own_rows = gidx.view(C, self.batch_real)[self.owned_t].reshape(-1)
            mean_real = ops.class_mean(gathered[own_rows].view(n_own, self.batch_real, D))
        elif video_sharded:
            # cheaper variant: (C, D) partial sums of this rank's videos -> all-reduce (0.4 MB); the summation order differs
            # from a single GPU's by re-association (1e-7 relative)
            sums = ops.class_sum_ragged(emb_real, offsets, C)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            # tensor divisor: IEEE division like the class_mean kernel (a Python scalar would become a reciprocal multiply)
            mean_real = (sums[self.owned_t] / torch.tensor(float(self.batch_real), device=self.device)).contiguous()
        else:
            mean_real = ops.class_mean(emb_real.view(n_own, self.batch_real, D))
        if joint:
            emb_syn = emb_syn.view(n_own, vpc, D)
        elif fused_syn:
            emb_syn = tc.embed_autograd(image_syn).view(n_own, vpc, D)
        else:
            emb_syn = net.embed(image_syn).view(n_own, vpc, D)
        self._mark('class means (+ all-gather)')
        loss = ops.dm_loss(mean_real, emb_syn)
        for p in (self.dynamic_syn, self.static_syn, *self.hal.parameters()):
            p.grad = None
        loss.backward()
        self._mark('loss + backward (dgrads, composer)')
        # ---- combine ranks: ONE small all-reduce [hallucinator grads (327) | loss]; the memory gradients are class-local
        loss_d = loss.detach().reshape(1).clone()
        allreduce_sum_([self.hal.encoder.weight.grad, self.hal.encoder.bias.grad, loss_d])
        self._mark('all-reduce [hal grad | loss]')
        # ---- optimizer steps (distill_s2d_ms.py:432-435), dense momentum SGD
        if self.train_static:
            self._sgd('static', self.static_syn, self.static_syn.grad, self.lr_static)
        self._sgd('dynamic', self.dynamic_syn, self.dynamic_syn.grad, self.lr_dynamic)
        self._sgd('hal_w', self.hal.encoder.weight, self.hal.encoder.weight.grad, self.lr_hal)
        self._sgd('hal_b', self.hal.encoder.bias, self.hal.encoder.bias.grad, self.lr_hal)
        self._mark('momentum SGD')
        self.last = dict(label=label, dynamic_idx=dynamic_idx, static_idx=static_idx, real_idx=real_idx,
                         emb_syn=emb_syn.detach(), mean_real=mean_real, image_syn=image_syn.detach())
        return loss_d[0]


class DMBaselineTrainer:
    """DM on leaf synthetic videos (distill_baseline.py:92-108, 334-356), SGD momentum 0.5."""

    def __init__(self, dataset, *, num_classes, channel=3, im_size=(112, 112), frames=16, ipc=1, batch_real=64,
                 lr_img=1.0, precision='bf16', image_syn=None, init='real', max_batch=128, device='cuda'):
        self.rank, self.world = _world()
        self.ds = dataset
        self.C, self.channel, self.im_size, self.frames = num_classes, channel, tuple(im_size), frames
        self.ipc, self.batch_real, self.lr_img = ipc, batch_real, lr_img
        self.device = torch.device(device)
        H, W = self.im_size
        if image_syn is None:
            image_syn = torch.randn(size=(num_classes * ipc, frames, channel, H, W), dtype=torch.float)
            if init == 'real':
                if self.world != 1:
                    raise RuntimeError("init='real' draws from every class; build image_syn on rank 0 and pass it in")
                for c in range(num_classes):                             # distill_baseline.py:96-100
                    image_syn[c * ipc:(c + 1) * ipc] = dataset.get_images(c, ipc).detach().cpu()
        self.image_syn = image_syn.detach().to(self.device).contiguous().requires_grad_(True)
        self._buf = None
        self.embedder = _RealEmbedder(frames, self.im_size, self.device, precision, max_batch)
        self.owned = owned_classes(num_classes, self.rank, self.world)
        self.last = {}

    def step(self, net=None, net_seed=None, real_idx=None):
        C, ipc = self.C, self.ipc
        if net is None:
            net = frozen_convnet3d(self.channel, C, self.im_size, self.frames, self.device, seed=net_seed)
        self.embedder.load(net)
        if real_idx is None:
            real_idx = self.ds.sample_all_classes(self.batch_real)
        own = self.owned
        n_own = len(own)
        rows = torch.as_tensor([c * ipc + i for c in own for i in range(ipc)], dtype=torch.long, device=self.device)
        emb_real = self.embedder(self.ds.videos, self.ds.local_index(real_idx[own]))
        D = emb_real.shape[1]
        mean_real = ops.class_mean(emb_real.view(n_own, self.batch_real, D))
        img = self.image_syn[rows] if self.world > 1 else self.image_syn
        if self.embedder.tc is not None:
            emb_syn = self.embedder.tc.embed_autograd(img).view(n_own, ipc, D)       # tensor cores, like the real branch
        else:
            emb_syn = net.embed(img).view(n_own, ipc, D)
        loss = ops.dm_loss(mean_real, emb_syn)
        self.image_syn.grad = None
        loss.backward()
        loss_d = loss.detach().reshape(1).clone()
        allreduce_sum_([self.image_syn.grad, loss_d])
        first = self._buf is None
        if first:
            self._buf = torch.empty_like(self.image_syn)
        ops.sgd_momentum_(self.image_syn.data, self.image_syn.grad.contiguous(), self._buf, self.lr_img, 0.5, first)
        self.last = dict(real_idx=real_idx, emb_syn=emb_syn.detach(), mean_real=mean_real)
        return loss_d[0]


class MTTS2DTrainer:
    """MTT + static/dynamic memory (distill_s2d_ms.py:197-300): unrolled ReparamModule student with
    second-order autograd through the conv trio.  precision='fp32': exact CUDA-core kernels (parity mode);
    'bf16': every fprop / dgrad / wgrad of the three feature convolutions (first and second order) runs as a
    tcgen05 GEMM with bf16 operands and fp32 accumulation (tc_trio.py)."""

    def __init__(self, *, num_classes, channel=3, im_size=(112, 112), frames=16, vpc=1, spc=2, dpc=2, syn_steps=10,
                 lr_dynamic=1e4, lr_hal=1e-2, lr_static=1e-4, lr_lr=1e-5, lr_teacher=0.01, train_static=False,
                 train_lr=True, batch_syn=None, static_syn=None, dynamic_syn=None, hal=None, device='cuda',
                 precision='fp32'):
        if precision not in ('fp32', 'bf16', 'bf16x3'):
            raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
        self.precision = precision
        self.C, self.channel, self.im_size, self.frames = num_classes, channel, tuple(im_size), frames
        self.vpc, self.spc, self.dpc, self.syn_steps = vpc, spc, dpc, syn_steps
        self.lr = dict(dynamic=lr_dynamic, hal=lr_hal, static=lr_static, lr=lr_lr)
        self.train_static, self.train_lr = train_static, train_lr
        self.batch_syn = batch_syn if batch_syn is not None else num_classes * vpc
        self.rank, self.world = _world()          # world > 1: every inner step's batch is sharded i % world == rank
        self.device = torch.device(device)
        H, W = self.im_size
        if static_syn is None:
            static_syn = torch.randn(size=(num_classes * spc, 3, H, W), dtype=torch.float)
        if dynamic_syn is None:
            dynamic_syn = torch.randn(size=(num_classes, dpc, frames, 1, H, W), dtype=torch.float)
        self.hal = (hal if hal is not None else Conv3DNet()).to(self.device)
        self.static_syn = static_syn.detach().to(self.device).contiguous().requires_grad_(train_static)
        self.dynamic_syn = dynamic_syn.detach().to(self.device).contiguous().requires_grad_(True)
        self.syn_lr = torch.tensor(lr_teacher).to(self.device).requires_grad_(train_lr)
        self._bufs = {}
        self.criterion = torch.nn.CrossEntropyLoss().to(self.device)
        self.last = {}

    def _sgd(self, name, p, lr, momentum):
        first = name not in self._bufs
        if first:
            self._bufs[name] = torch.empty_like(p)
        ops.sgd_momentum_(p.data, p.grad.contiguous(), self._bufs[name], lr, momentum, first)

    def step(self, start_params, target_params, student_net=None, net_seed=None):
        """start_params / target_params: lists of expert tensors (buffer.py layout) or flat tensors."""
        from . import tc_trio
        prev = ops.set_conv_backend(ops.backend_for_precision(self.precision))
        tc_trio.xcol_cache_begin()               # the im2col of a saved activation serves both wgrads of the iteration
        try:
            return self._step(start_params, target_params, student_net, net_seed)
        finally:
            tc_trio.xcol_cache_end()
            ops.set_conv_backend(prev)

    def _step(self, start_params, target_params, student_net=None, net_seed=None):
        C, vpc, spc, dev = self.C, self.vpc, self.spc, self.device
        if student_net is None:
            if net_seed is not None:
                torch.random.manual_seed(int(net_seed))
                base = ConvNet3D(self.channel, C, 128, 3, 'relu', 'none', 'maxpooling', self.frames, self.im_size).to(dev)
            else:
                base = get_network('ConvNet3D', self.channel, C, self.im_size, frames=self.frames, dist=False).to(dev)
            student_net = ReparamModule(base)
        student_net.train()
        num_params = student_net.param_numel

        def flat(ps):
            if torch.is_tensor(ps):
                return ps.to(dev).reshape(-1)
            return torch.cat([p.data.to(dev).reshape(-1) for p in ps], 0)
        target = flat(target_params)
        starting = flat(start_params)
        student_params = [starting.clone().requires_grad_(True)]
        chunks, draws = [], []
        for _ in range(self.syn_steps):                                     # distill_s2d_ms.py:238-266
            if not chunks:
                indices = torch.randperm(C * vpc, device=dev)
                chunks = list(torch.split(indices, self.batch_syn))
            these = chunks.pop()
            label = these // vpc
            idx = these % vpc
            dynamic_idx = 2 * idx + torch.randint(2, (these.shape[0],), device=dev)
            static_idx = spc * label + 2 * idx + torch.randint(2, (these.shape[0],), device=dev)
            n_step = int(these.shape[0])
            sl = slice(self.rank, None, self.world)               # this rank's videos of the step batch (all for 1 rank)

            def shard_loss(theta, label=label, dynamic_idx=dynamic_idx, static_idx=static_idx, sl=sl, n_step=n_step):
                if label[sl].numel() == 0:                      # fewer videos in this step batch than ranks: this rank contributes 0
                    return theta.sum() * 0.0                    # (still attached to theta, so the collectives of the backward line up)
                x = self.hal.compose(self.static_syn, self.dynamic_syn, static_idx[sl], label[sl], dynamic_idx[sl])
                out = student_net(x, flat_param=theta)
                # CrossEntropyLoss() is the batch mean (distill_s2d_ms.py:262): the shard contributes sum / B
                return torch.nn.functional.cross_entropy(out, label[sl].long(), reduction='sum') / n_step
            grad = sharded_inner_grad(student_params[-1], shard_loss, self.world)
            student_params.append(student_params[-1] - self.syn_lr * grad)
            draws.append((these, dynamic_idx, static_idx))
        param_loss = torch.nn.functional.mse_loss(student_params[-1], target, reduction='sum') / num_params
        param_dist = torch.nn.functional.mse_loss(starting, target, reduction='sum') / num_params
        grand_loss = param_loss / param_dist
        for p in (self.dynamic_syn, self.static_syn, self.syn_lr, *self.hal.parameters()):
            p.grad = None
        grand_loss.backward()
        if self.world > 1:
            # the grand loss is replicated; each rank holds the gradient that flowed through ITS shard of every step
            grads = [self.dynamic_syn.grad, self.hal.encoder.weight.grad, self.hal.encoder.bias.grad]
            if self.train_static:
                grads.append(self.static_syn.grad)
            allreduce_sum_(grads)
        if self.train_static:
            self._sgd('static', self.static_syn, self.lr['static'], 0.95)
        self._sgd('dynamic', self.dynamic_syn, self.lr['dynamic'], 0.95)
        self._sgd('hal_w', self.hal.encoder.weight, self.lr['hal'], 0.95)
        self._sgd('hal_b', self.hal.encoder.bias, self.lr['hal'], 0.95)
        if self.train_lr:
            g = self.syn_lr.grad.reshape(1).contiguous()
            p = self.syn_lr.data.reshape(1)
            first = 'lr' not in self._bufs
            if first:
                self._bufs['lr'] = torch.empty_like(p)
            # a 1-element tensor is below the kernel's 16-byte alignment contract: plain torch here
            buf = self._bufs['lr']
            buf.copy_(g) if first else buf.mul_(0.9).add_(g)
            p.sub_(self.lr['lr'] * buf)
            self.syn_lr.data = self.syn_lr.data.clip(min=0.001)
        self.last = dict(draws=draws, param_loss=param_loss.detach(), param_dist=param_dist.detach())
        return grand_loss.detach()


class MTTBaselineTrainer:
    """MTT on leaf synthetic videos (distill_baseline.py:92-108, 196-300): the unrolled ReparamModule student of
    MTTS2DTrainer without the composer — ``x = image_syn[these]``, SGD(momentum=0.5) on the videos and on syn_lr.
    The step batches come from ``torch.randperm(len(image_syn))`` on the HOST generator, as in the reference
    (:235), so a seeded run reproduces its index stream."""

    def __init__(self, *, num_classes, channel=3, im_size=(112, 112), frames=16, ipc=1, syn_steps=10, lr_img=1.0,
                 lr_lr=1e-5, lr_teacher=0.001, train_lr=False, batch_syn=None, image_syn=None, device='cuda',
                 precision='fp32'):
        if precision not in ('fp32', 'bf16', 'bf16x3'):
            raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
        self.precision = precision
        self.C, self.channel, self.im_size, self.frames = num_classes, channel, tuple(im_size), frames
        self.ipc, self.syn_steps, self.lr_img, self.lr_lr, self.train_lr = ipc, syn_steps, lr_img, lr_lr, train_lr
        self.batch_syn = batch_syn if batch_syn is not None else num_classes * ipc
        self.rank, self.world = _world()
        self.device = torch.device(device)
        H, W = self.im_size
        if image_syn is None:
            image_syn = torch.randn(size=(num_classes * ipc, frames, channel, H, W), dtype=torch.float)
        self.image_syn = image_syn.detach().to(self.device).contiguous().requires_grad_(True)
        self.label_syn = torch.tensor(np.stack([np.ones(ipc) * i for i in range(0, num_classes)]), dtype=torch.long,
                                      requires_grad=False, device=self.device).view(-1)
        self.syn_lr = torch.tensor(lr_teacher).to(self.device).requires_grad_(train_lr)
        self._bufs = {}
        self.last = {}

    def step(self, start_params, target_params, student_net=None, net_seed=None):
        from . import tc_trio
        prev = ops.set_conv_backend(ops.backend_for_precision(self.precision))
        tc_trio.xcol_cache_begin()               # the im2col of a saved activation serves both wgrads of the iteration
        try:
            return self._step(start_params, target_params, student_net, net_seed)
        finally:
            tc_trio.xcol_cache_end()
            ops.set_conv_backend(prev)

    def _step(self, start_params, target_params, student_net, net_seed):
        C, dev = self.C, self.device
        if student_net is None:
            if net_seed is not None:
                torch.random.manual_seed(int(net_seed))
                base = ConvNet3D(self.channel, C, 128, 3, 'relu', 'none', 'maxpooling', self.frames, self.im_size).to(dev)
            else:
                base = get_network('ConvNet3D', self.channel, C, self.im_size, frames=self.frames, dist=False).to(dev)
            student_net = ReparamModule(base)
        student_net.train()
        num_params = student_net.param_numel

        def flat(ps):
            if torch.is_tensor(ps):
                return ps.to(dev).reshape(-1)
            return torch.cat([p.data.to(dev).reshape(-1) for p in ps], 0)
        target = flat(target_params)
        starting = flat(start_params)
        student_params = [starting.clone().requires_grad_(True)]
        chunks, draws = [], []
        for _ in range(self.syn_steps):                                      # distill_baseline.py:231-252
            if not chunks:
                indices = torch.randperm(len(self.image_syn))               # host generator, like the reference
                chunks = list(torch.split(indices, self.batch_syn))
            these = chunks.pop().to(dev)
            n_step = int(these.shape[0])
            sl = slice(self.rank, None, self.world)

            def shard_loss(theta, these=these, sl=sl, n_step=n_step):
                if these[sl].numel() == 0:
                    return theta.sum() * 0.0
                out = student_net(self.image_syn[these[sl]], flat_param=theta)
                return torch.nn.functional.cross_entropy(out, self.label_syn[these[sl]], reduction='sum') / n_step
            grad = sharded_inner_grad(student_params[-1], shard_loss, self.world)
            student_params.append(student_params[-1] - self.syn_lr * grad)
            draws.append(these)
        param_loss = torch.nn.functional.mse_loss(student_params[-1], target, reduction='sum') / num_params
        param_dist = torch.nn.functional.mse_loss(starting, target, reduction='sum') / num_params
        grand_loss = param_loss / param_dist
        self.image_syn.grad = None
        self.syn_lr.grad = None
        grand_loss.backward()
        if self.world > 1:
            allreduce_sum_([self.image_syn.grad])
        first = 'img' not in self._bufs
        if first:
            self._bufs['img'] = torch.empty_like(self.image_syn)
        ops.sgd_momentum_(self.image_syn.data, self.image_syn.grad.contiguous(), self._bufs['img'], self.lr_img, 0.5, first)
        if self.train_lr:
            g = self.syn_lr.grad.reshape(1)
            p = self.syn_lr.data.reshape(1)
            first = 'lr' not in self._bufs
            if first:
                self._bufs['lr'] = torch.empty_like(p)
            buf = self._bufs['lr']
            buf.copy_(g) if first else buf.mul_(0.5).add_(g)
            p.sub_(self.lr_lr * buf)
            self.syn_lr.data = self.syn_lr.data.clip(min=0.001)
        self.last = dict(draws=draws, param_loss=param_loss.detach(), param_dist=param_dist.detach())
        return grand_loss.detach()
