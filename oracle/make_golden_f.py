#!/usr/bin/env python
"""Golden vectors for the (f) rows of SURVEY.md section 8 — callers of the hot path — from the LIVE reference.

Run in the build container only (needs /root/reference):   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_f.py

* f1  utils.epoch('train') and utils.evaluate_synset (utils.py:752-886) on a seeded ConvNet3D and hash-generated videos;
* f3  the expert loop of buffer.py:65-98 (get_network under a patched clock, SGD, one `epoch('train')` per epoch, parameter
      snapshot after every epoch);
* f4  distill_coreset.py:75-110: herding / k-center indices selected from the reference net's own embeddings.
The reference modules are executed unmodified on the CPU; the only intervention is Dropout(p=0) (the CPU and CUDA random
streams cannot be matched) and batch sizes >= the dataset so that the DataLoader shuffle has no effect.  Outputs of the
REFERENCE are stored in tests/golden/f_rows.npz; inputs are regenerated from hashes (oracle/synth.py).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, REPO, GOLD  # noqa: E402

ref_networks, ref_utils, ref_reparam = import_reference()
sys.path.insert(0, REPO)
from oracle import synth  # noqa: E402

C, T, H = 4, 8, 64
N_TRAIN, N_TEST = 3, 2


def data():
    xtr = synth.hash_uniform((C * N_TRAIN, T, 3, H, H), 801)
    ytr = torch.arange(C).repeat_interleave(N_TRAIN)
    xte = synth.hash_uniform((C * N_TEST, T, 3, H, H), 802)
    yte = torch.arange(C).repeat_interleave(N_TEST)
    return xtr, ytr, xte, yte


def make_args(**kw):
    a = type('A', (), {})()
    a.device, a.model, a.eval_mode = 'cpu', 'ConvNet3D', 'S'
    a.lr_net, a.epoch_eval_train, a.batch_train = 0.01, 2, C * N_TRAIN
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def ref_net(seed):
    net = ref_networks.ConvNet3D(channel=3, num_classes=C, net_width=128, net_depth=3, net_act='relu', net_norm='none',
                                 net_pooling='maxpooling', im_size=(H, H), frames=T)
    net.load_state_dict(synth.synth_convnet3d_params(seed, num_classes=C))
    net.dropout.p = 0.0
    return net


def summary(out, tag, net):
    for name, p in net.state_dict().items():
        s, samp = synth.summarize(p, stride=53)
        out[f'{tag}.{name}.sums'] = s
        out[f'{tag}.{name}.sample'] = samp


def main():
    torch.set_num_threads(8)
    out = {}
    xtr, ytr, xte, yte = data()
    testloader = torch.utils.data.DataLoader(ref_utils.TensorDataset(xte, yte), batch_size=64, shuffle=False, num_workers=0)
    args = make_args()
    # ---- f1a: one training epoch
    net = ref_net(70)
    trainloader = torch.utils.data.DataLoader(ref_utils.TensorDataset(xtr, ytr), batch_size=args.batch_train, shuffle=True, num_workers=0)
    optim = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0005)
    loss, acc, per = ref_utils.epoch('train', trainloader, net, optim, nn.CrossEntropyLoss(), args)
    out['epoch_train.loss'], out['epoch_train.acc'] = np.float64(loss), np.float64(acc)
    summary(out, 'epoch_train', net)
    with torch.no_grad():
        loss_t, acc_t, per_t = ref_utils.epoch('test', testloader, net, None, nn.CrossEntropyLoss(), args)
    out['epoch_test.loss'], out['epoch_test.acc'] = np.float64(loss_t), np.float64(acc_t)
    out['epoch_test.per_class'] = np.asarray([np.nan if v is None else v for v in per_t], dtype=np.float64)
    # ---- f1b: evaluate_synset, Epoch = 2 (three training epochs, lr x0.1 after epoch 2, test at the end)
    net = ref_net(71)
    _, acc_train, acc_test, acc_per = ref_utils.evaluate_synset(0, net, xtr, ytr, testloader, args, mode='none')
    out['evaluate_synset.acc_train'], out['evaluate_synset.acc_test'] = np.float64(acc_train), np.float64(acc_test)
    out['evaluate_synset.acc_per'] = np.asarray([np.nan if v is None else v for v in acc_per], dtype=np.float64)
    summary(out, 'evaluate_synset', net)
    # ---- f3: buffer.py expert loop, 2 epochs, lr_teacher 0.01 mom 0 l2 0 (defaults of buffer.py:113-124), no decay
    seed = 4242
    real_time = time.time
    ref_utils.time.time = lambda: seed / 1000.0 + 1e-7
    try:
        teacher = ref_utils.get_network('ConvNet3D', 3, C, (H, H), frames=T, dist=False)
    finally:
        ref_utils.time.time = real_time
    teacher.dropout.p = 0.0
    teacher.train()
    bargs = make_args(lr_teacher=0.01, mom=0.0, l2=0.0, train_epochs=2)
    optim = torch.optim.SGD(teacher.parameters(), lr=bargs.lr_teacher, momentum=bargs.mom, weight_decay=bargs.l2)
    optim.zero_grad()
    out['buffer.seed'] = np.int64(seed)
    for name, p in teacher.state_dict().items():
        out[f'buffer.e0.{name}.sums'] = synth.summarize(p, stride=53)[0]
    for e in range(bargs.train_epochs):
        tl, ta, _ = ref_utils.epoch('train', dataloader=trainloader, net=teacher, optimizer=optim, criterion=nn.CrossEntropyLoss(), args=bargs)
        out[f'buffer.e{e + 1}.train_loss'], out[f'buffer.e{e + 1}.train_acc'] = np.float64(tl), np.float64(ta)
        summary(out, f'buffer.e{e + 1}', teacher)
    # ---- f4: coreset indices from the reference net's embeddings (eval mode, distill_coreset.py:60-110)
    import textwrap
    src = open(os.path.join(os.environ.get('VD_REFERENCE', '/root/reference'), 'distill_coreset.py')).read()
    net = ref_net(72)
    net.eval()
    pool = synth.hash_uniform((C * 6, T, 3, H, H), 803)
    plab = torch.arange(C).repeat_interleave(6)
    for method, result, ipc in (('k-center', 'idx_centers', 1), ('herding', 'idx_selected', 2)):
        seg = src[src.index("args.method == '%s'" % method):]
        seg = seg[seg.index('features = embed(imgs)') + len('features = embed(imgs)'):]
        seg = seg[:seg.index('image_syn[c*args.ipc')]
        code = textwrap.dedent('\n'.join(seg.splitlines()[1:]))
        chosen = []
        for c in range(C):
            idx = torch.nonzero(plab == c).flatten()
            with torch.no_grad():
                features = net.embed(pool[idx])
            a = type('A', (), {})()
            a.ipc = ipc
            ns = {'torch': torch, 'np': np, 'features': features, 'args': a}
            exec(code, ns)
            chosen += [int(idx[j]) for j in ns[result]]
        out[f'coreset.{method}'] = np.asarray(chosen, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, 'f_rows.npz'), **out)
    print('f rows ok:', {k: (float(v) if np.ndim(v) == 0 else v.tolist()) for k, v in out.items()
                         if k.endswith(('loss', 'acc', 'acc_train', 'acc_test')) or k.startswith('coreset.') and 'margin' not in k})


if __name__ == '__main__':
    main()
