"""Drop-in counterparts of the hot-path helpers of /root/reference/utils.py:

get_network (:518-625, ConvNet3D branch), Conv3DNet (:1178-1197), TensorDataset /
MultiStaticSharedDataset (:462-508), epoch / evaluate_synset (:752-886) — same names,
arguments and return values; tensor work goes through libvd_b200.
"""
import random
import time

import torch
import torch.nn as nn
from torch.utils.data import Dataset

from . import ops
from .networks import ConvNet3D


def get_default_convnet_setting():
    net_width, net_depth, net_act, net_norm, net_pooling = 128, 3, 'relu', 'instancenorm', 'avgpooling'
    return net_width, net_depth, net_act, net_norm, net_pooling


def get_network(model, channel, num_classes, im_size=(32, 32), frames=16, dist=True):
    """Reference factory (utils.py:518-625) restricted to the model the distillation scripts use.

    Like the reference it reseeds torch's global generator from the wall clock (:519) and
    overrides the default setting with net_norm='none', net_pooling='maxpooling' (:608-609).
    ``dist=True`` moves the net to the GPU; nn.DataParallel is NOT used — multi-GPU runs are one
    process per GPU (see video_distillation_b200.distributed).
    """
    torch.random.manual_seed(int(time.time() * 1000) % 100000)
    net_width, net_depth, net_act, net_norm, net_pooling = get_default_convnet_setting()
    if model == 'ConvNet3D':
        net = ConvNet3D(channel=channel, num_classes=num_classes, net_width=net_width, net_depth=net_depth,
                        net_act=net_act, net_norm='none', net_pooling='maxpooling', im_size=im_size, frames=frames)
    else:
        raise NotImplementedError(
            'get_network (B200 hot path) only builds ConvNet3D, the model every distillation script selects; got %s' % model)
    if dist:
        if not torch.cuda.is_available():
            raise RuntimeError('video_distillation_b200 needs a CUDA device (no CPU fallback)')
        net = net.to('cuda')
    return net


def get_time():
    return str(time.strftime("[%Y-%m-%d %H:%M:%S]", time.localtime()))


def get_eval_pool(eval_mode, model, model_eval):
    return [model_eval]


class Conv3DNet(nn.Module):
    """The hallucinator (utils.py:1178-1197): a learnable Conv3d(4->3, k=3, pad=1) over
    cat([static broadcast over T, dynamic]).  ``forward(static, dynamic)`` keeps the reference
    signature (already gathered rows); ``compose`` is the fused form that takes the memories and
    the index vectors and folds the gathers (distill_s2d_ms.py:409-410) into the kernel."""

    def __init__(self, in_channel=4, mid_channel=3, out_channel=3, img_size=112, kernel_size=3, mode='concat'):
        super().__init__()
        if mode != 'concat' or in_channel != 4 or mid_channel != 3 or kernel_size != 3:
            raise NotImplementedError('Conv3DNet (B200): only the default concat 4->3, k=3 configuration is supported')
        self.mode = mode
        self.encoder = nn.Conv3d(in_channel, mid_channel, kernel_size, padding=1)   # parameters + default init only

    def compose(self, static_syn, dynamic_syn, static_idx, label, dynamic_idx, unique_rows=False):
        return ops.compose(static_syn, dynamic_syn, self.encoder.weight, self.encoder.bias, static_idx, label, dynamic_idx, unique_rows)

    def forward(self, static, dynamic):
        b = dynamic.shape[0]
        ar = torch.arange(b, device=dynamic.device)
        return ops.compose(static, dynamic.unsqueeze(1), self.encoder.weight, self.encoder.bias,
                           ar, ar, torch.zeros_like(ar))


class TensorDataset(Dataset):
    def __init__(self, images, labels):
        self.images = images.detach().float()
        self.labels = labels.detach()

    def __getitem__(self, index):
        return self.images[index], self.labels[index]

    def __len__(self):
        return self.images.shape[0]


class MultiStaticSharedDataset(Dataset):
    """Eval-time view of (static, dynamic, hallucinators) — utils.py:462-496 index pairing."""

    def __init__(self, static, dynamic, hallucinator):
        self.static = static.detach().float()
        self.dynamic = dynamic.detach().float()
        self.hallucinator = hallucinator
        self.n_s = static.shape[0]
        self.n_c, self.dpc = dynamic.shape[0], dynamic.shape[1]

    def pick(self, index):
        """(static row, label, dynamic memory, hallucinator) of sample ``index`` — the reference's pairing and its order of
        ``random.randint`` draws (utils.py:470-484)."""
        per_s = self.n_s // self.n_c
        if per_s == 10:
            label, idx = index // 5, index % 5
            static_idx = label * per_s + 2 * idx + random.randint(0, 1)
            dynamic_idx = 2 * idx + random.randint(0, 1)
        elif per_s == 2:
            label = index
            static_idx = random.randint(0, per_s - 1) + label * per_s
            dynamic_idx = random.randint(0, self.dpc - 1)
        else:
            raise ValueError('MultiStaticSharedDataset: spc must be 2 (vpc=1) or 10 (vpc=5)')
        hal_idx = random.randint(0, len(self.hallucinator) - 1)
        return static_idx, label, dynamic_idx, hal_idx

    def __getitem__(self, index):
        static_idx, label, dynamic_idx, hal_idx = self.pick(index)
        hal = self.hallucinator[hal_idx]
        dev = self.dynamic.device
        with torch.no_grad():
            video = hal.compose(self.static, self.dynamic, torch.tensor([static_idx], device=dev),
                                torch.tensor([label], device=dev), torch.tensor([dynamic_idx], device=dev))
        return video[0], label

    def __len__(self):
        if self.n_s == self.n_c * 10:
            return self.n_c * 5
        if self.n_s == self.n_c * 2:
            return self.n_c
        raise ValueError('MultiStaticSharedDataset: spc must be 2 (vpc=1) or 10 (vpc=5)')


class EpochStats:
    """Running statistics of ``epoch`` kept ON THE DEVICE: the reference converts logits and labels to numpy after every
    batch (utils.py:775-781, 804-816), i.e. one host synchronisation per batch; here the counts accumulate in a few small
    tensors and are read back once per epoch.  Same definitions: accuracy = first-argmax match, top-k = label among the k
    largest logits (all classes when there are fewer than k), per-class accuracy over the samples seen."""

    def __init__(self, num_classes, device):
        self.C = num_classes
        self.sums = torch.zeros(6, dtype=torch.float64, device=device)           # loss*n, matched, top1, top3, top5, n
        self.cls = torch.zeros(2, num_classes, dtype=torch.float64, device=device)   # per class: correct, seen

    @torch.no_grad()
    def add(self, output, lab, loss, train):
        n_b = lab.shape[0]
        matched = output.argmax(dim=-1).eq(lab)
        k5 = min(5, output.shape[1])
        top = output.topk(k5, dim=-1).indices.eq(lab[:, None])
        zero = torch.zeros((), dtype=torch.float64, device=output.device)
        vals = [loss.detach().double() * n_b, matched.sum().double(),
                zero if train else top[:, :1].any(1).sum().double(),              # the train branch of the reference only counts top-5
                zero if train else top[:, :min(3, k5)].any(1).sum().double(),
                top.any(1).sum().double(), zero + n_b]
        self.sums += torch.stack(vals)
        self.cls[0] += torch.bincount(lab, weights=matched.double(), minlength=self.C)[:self.C]
        self.cls[1] += torch.bincount(lab, minlength=self.C)[:self.C].double()

    def result(self, top5_mode):
        sums = self.sums.tolist()                                                  # the one device -> host read of the epoch
        correct, seen = self.cls.tolist()
        n = sums[5]
        loss_avg, acc_avg = sums[0] / n, sums[1] / n
        present = [i for i in range(self.C) if seen[i] > 0]
        # the reference builds a dict over the classes seen and lists range(len(dict)) of it (utils.py:834-839)
        per_class = [correct[i] / seen[i] if seen[i] > 0 else None for i in range(len(present))]
        if top5_mode:
            return loss_avg, [acc_avg, sums[2] / n, sums[3] / n, sums[4] / n], per_class
        return loss_avg, acc_avg, per_class


def epoch(mode, dataloader, net, optimizer, criterion, args):
    """utils.py:752-845: one training epoch, or three test passes; returns (loss, acc, acc_per_class)."""
    net = net.to(args.device)
    net.train() if mode == 'train' else net.eval()
    stats = None
    for _ in range(1 if mode == 'train' else 3):
        for datum in dataloader:
            img = datum[0].float().to(args.device)
            img = (img - img.mean()) / img.std()
            lab = datum[1].long().to(args.device)
            output = net(img)
            loss = criterion(output, lab)
            if stats is None:
                stats = EpochStats(output.shape[1], output.device)
            stats.add(output, lab, loss, mode == 'train')
            if mode == 'train':
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
    return stats.result(getattr(args, 'eval_mode', None) == 'top5')


def evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, mode='hallucinator',
                    return_loss=False, test_freq=None):
    """utils.py:848-886: train ``net`` on the synthetic set for args.epoch_eval_train epochs
    (SGD m=0.9 wd=5e-4, lr x0.1 after Epoch//2+1) and test; returns (net, acc_train, acc_test, acc_per)."""
    # args.precision = 'bf16' (optional, not a reference flag): train / test on the tensor-core conv trio
    prev_backend = ops.set_conv_backend({'bf16': 'tc', 'bf16x3': 'tc_x3'}.get(getattr(args, 'precision', 'fp32'), 'fp32'))
    try:
        return _evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, mode, return_loss, test_freq)
    finally:
        ops.set_conv_backend(prev_backend)


def _evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, mode, return_loss, test_freq):
    lr = float(args.lr_net)
    Epoch = int(args.epoch_eval_train)
    lr_schedule = [Epoch // 2 + 1]
    optimizer = torch.optim.SGD(net.parameters(), lr=lr, momentum=0.9, weight_decay=0.0005)
    criterion = nn.CrossEntropyLoss().to(args.device)
    if mode == 'none':
        dst_train = TensorDataset(images_train, labels_train)
    elif mode == 'multi-static':
        dst_train = MultiStaticSharedDataset(images_train[0], images_train[1], images_train[2])
    else:
        raise NotImplementedError
    trainloader = torch.utils.data.DataLoader(dst_train, batch_size=args.batch_train, shuffle=True, num_workers=0)
    start = time.time()
    acc_test, acc_per, loss_train, acc_train = 0.0, None, 0.0, 0.0
    for ep in range(Epoch + 1):
        loss_train, acc_train, _ = epoch('train', trainloader, net, optimizer, criterion, args)
        if (test_freq is None and ep == Epoch) or (test_freq is not None and ep % test_freq == 0 and ep != 0):
            with torch.no_grad():
                loss_test, acc_test, acc_per = epoch('test', testloader, net, optimizer, criterion, args)
        if ep in lr_schedule:
            lr *= 0.1
            optimizer = torch.optim.SGD(net.parameters(), lr=lr, momentum=0.9, weight_decay=0.0005)
    if getattr(args, 'eval_mode', None) != 'top5':
        print('%s Evaluate_%02d: Ep %d time = %ds loss = %.6f train acc = %.2f, test acc = %.2f' % (
            get_time(), it_eval, Epoch, int(time.time() - start), loss_train, acc_train * 100, acc_test * 100))
    return net, acc_train, acc_test, acc_per


def get_loops(ipc, dataset=None):
    """utils.py:691-709 (outer / inner loop counts of the DC baselines; the distillation drivers only print them)."""
    table = {1: (1, 1), 5: (1, 1), 10: (10, 50), 20: (20, 25), 30: (30, 20), 40: (40, 15), 50: (50, 10)}
    if ipc not in table:
        raise SystemExit('loop hyper-parameters are not defined for %d ipc' % ipc)
    return table[ipc]


class ParamDiffAug:
    """utils.py:999-1009: parameter bag of the differentiable augmentation (unused by the video scripts' DM / MTT paths)."""

    def __init__(self):
        self.aug_mode = 'S'
        self.prob_flip = 0.5
        self.ratio_scale = 1.2
        self.ratio_rotate = 15.0
        self.ratio_crop_pad = 0.125
        self.ratio_cutout = 0.5
        self.brightness = 1.0
        self.saturation = 2.0
        self.contrast = 0.5


def _off_path(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(f'{name} belongs to the DC / DSA baselines of the reference, which are outside the B200 hot path '
                                  f'(SURVEY.md §2); the DM and MTT paths never call it')
    fn.__name__ = name
    return fn


DiffAugment = _off_path('DiffAugment')
match_loss = _off_path('match_loss')
get_daparam = _off_path('get_daparam')


def get_dataset(dataset, data_path, num_workers=0, img_size=(112, 112), split_num=1, split_id=0, split_mode='mean', **kw):
    """utils.py:21 (same argument order) — see video_distillation_b200/datasets.py for what is supported."""
    from .datasets import get_dataset as _get
    return _get(dataset, data_path, num_workers, img_size, split_num, split_id, split_mode, **kw)
