"""Host-side logic that needs no GPU: drop-in module surface, ReparamModule flattening, sampling."""
import inspect

import numpy as np
import pytest
import torch

import oracle


def test_dropin_module_names_and_signatures():
    import networks
    import reparam_module
    import utils
    sig = inspect.signature(utils.get_network)
    assert list(sig.parameters) == ['model', 'channel', 'num_classes', 'im_size', 'frames', 'dist']
    sig = inspect.signature(utils.evaluate_synset)
    assert list(sig.parameters) == ['it_eval', 'net', 'images_train', 'labels_train', 'testloader', 'args', 'mode',
                                    'return_loss', 'test_freq']
    sig = inspect.signature(networks.ConvNet3D.__init__)
    assert list(sig.parameters)[1:] == ['channel', 'num_classes', 'net_width', 'net_depth', 'net_act', 'net_norm',
                                        'net_pooling', 'frames', 'im_size', 'dropout_keep_prob']
    sig = inspect.signature(utils.Conv3DNet.__init__)
    assert list(sig.parameters)[1:] == ['in_channel', 'mid_channel', 'out_channel', 'img_size', 'kernel_size', 'mode']
    assert hasattr(reparam_module.ReparamModule, 'embed') and hasattr(reparam_module.ReparamModule, 'forward')


def test_get_network_seeding_and_param_layout(monkeypatch):
    """get_network reseeds from the clock (utils.py:519); with the clock patched the parameters equal the
    oracle's seeded init bit for bit, in the reference's state_dict order."""
    from video_distillation_b200 import utils
    monkeypatch.setattr(utils.time, 'time', lambda: 4321 / 1000.0 + 1e-7)
    net = utils.get_network('ConvNet3D', 3, 50, (112, 112), dist=False)
    ref = oracle.init_convnet3d(4321, 3, 50)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(torch.equal(sd[k], ref[k]) for k in ref)
    mods = [type(m).__name__ for m in net.features]
    assert mods == ['Conv3d', 'ReLU', 'MaxPool3d'] * 3                      # none / maxpooling (utils.py:608-609)
    assert net.features[2].kernel_size == (1, 2, 2) and net.features[5].kernel_size == (2, 2, 2)


def test_reparam_module_flat_layout():
    from video_distillation_b200.networks import ConvNet3D
    from video_distillation_b200.reparam_module import ReparamModule
    torch.manual_seed(0)
    net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', 16, (112, 112))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    rp = ReparamModule(net)
    assert rp.param_numel == 3647666 and [n for n, _ in rp.named_parameters()] == ['flat_param']
    assert torch.equal(rp.flat_param.detach(), oracle.flatten_params(sd))
    assert rp._param_numels == (28224, 64, 1204224, 128, 2408448, 128, 6400, 50)
    # views are re-installed around a call and restored afterwards
    other = torch.zeros_like(rp.flat_param)
    with rp.unflattened_param(other):
        assert rp.module.features[0].weight.data_ptr() == other.data_ptr()
    assert rp.module.features[0].weight.data_ptr() == rp.flat_param.data_ptr()


def test_device_dataset_sampling_is_reference_stream():
    """Every rank replays the full numpy stream; shards only differ in what they keep."""
    from video_distillation_b200.distill import DeviceDataset, owned_classes
    C, per = 5, 6
    videos = torch.zeros(C * per, 1, 3, 2, 2)
    labels = [c for c in range(C) for _ in range(per)]
    indices_class = [list(range(c * per, (c + 1) * per)) for c in range(C)]
    np.random.seed(3)
    want = np.stack([oracle.sample_real_indices(indices_class, c, 4) for c in range(C)])
    for rank in range(2):
        ds = DeviceDataset(videos, labels, C, 'cpu', rank=rank, world=2)
        np.random.seed(3)
        got = ds.sample_all_classes(4)
        assert np.array_equal(got, want)
        own = owned_classes(C, rank, 2)
        loc = ds.local_index(got[own])
        assert ds.videos.shape[0] == len(own) * per and int(loc.max()) < ds.videos.shape[0]
    assert owned_classes(50, 0, 8) == [0, 8, 16, 24, 32, 40, 48] and len(owned_classes(50, 7, 8)) == 6


def test_multistatic_dataset_pairing_matches_the_live_reference():
    """tests/golden/multistatic.npz: (index, static row, label, dynamic memory) chosen by the reference's
    MultiStaticSharedDataset.__getitem__ (recording hallucinator, random.seed(21)) — same pairing, same consumption of random."""
    import os
    import random
    import numpy as np
    import torch
    from video_distillation_b200.utils import MultiStaticSharedDataset
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'multistatic.npz'))
    for tag, C, spc, dpc in (('vpc1', 4, 2, 2), ('vpc5', 3, 10, 10)):
        ds = MultiStaticSharedDataset(torch.zeros(C * spc, 3, 2, 2), torch.zeros(C, dpc, 2, 1, 2, 2), torch.nn.ModuleList([torch.nn.Identity()] * 3))
        assert len(ds) == int(gold[tag + '_len'])
        random.seed(21)
        rows = []
        for index in list(range(len(ds))) * 2:
            static_idx, label, dynamic_idx, hal_idx = ds.pick(index)
            assert 0 <= hal_idx < 3
            rows.append((index, static_idx, label, dynamic_idx))
        assert np.array_equal(np.asarray(rows, dtype=np.int64), gold[tag])
        assert random.random() == float(gold[tag + '_next_random'])


@pytest.mark.parametrize('tag,test_freq', [('final', None), ('freq3', 3)])
def test_evaluate_synset_schedule_matches_the_live_reference(tag, test_freq, monkeypatch):
    """tests/golden/eval_schedule.npz: the (train / test, lr, momentum, weight decay) sequence of epoch() calls the reference's
    evaluate_synset makes for Epoch = 7 (recorded by oracle/make_golden.py with utils.epoch patched)."""
    import os
    import numpy as np
    import torch
    from video_distillation_b200 import utils as U
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'eval_schedule.npz'))
    calls = []

    def recorder(mode, dataloader, net, optimizer, criterion, args):
        g = optimizer.param_groups[0]
        calls.append((0 if mode == 'train' else 1, g['lr'], g['momentum'], g['weight_decay']))
        return 0.5, 0.25, [0.25]
    monkeypatch.setattr(U, 'epoch', recorder)
    args = type('A', (), {})()
    args.lr_net, args.epoch_eval_train, args.device, args.batch_train, args.eval_mode = 0.02, 7, 'cpu', 4, 'S'
    _, acc_train, acc_test, _ = U.evaluate_synset(0, torch.nn.Linear(3, 2), torch.zeros(6, 2, 3, 4, 4), torch.zeros(6).long(), None, args,
                                                  mode='none', test_freq=test_freq)
    assert np.allclose(np.asarray(calls, dtype=np.float64), gold[tag], rtol=0, atol=1e-15)
    assert [acc_train, acc_test] == gold[tag + '_ret'].tolist()


def test_video_sharded_dataset_covers_every_draw_exactly_once():
    """DeviceDataset(shard='video'): every class is spread over the ranks by position; the ranks' local shares of a draw
    partition it, keep the sampled order inside a class, and their per-class partial sums add up to the class sums."""
    import numpy as np
    import torch
    from video_distillation_b200.distill import DeviceDataset
    C, per, world, n = 5, 12, 4, 9
    labels = [c for c in range(C) for _ in range(per)]
    rng = np.random.RandomState(3)
    order = rng.permutation(len(labels))                      # classes interleaved in dataset order
    labels = [labels[i] for i in order]
    videos = torch.arange(len(labels), dtype=torch.float32).view(-1, 1, 1, 1, 1).expand(-1, 2, 3, 4, 4).contiguous()
    shards = [DeviceDataset(videos, labels, C, 'cpu', r, world, shard='video') for r in range(world)]
    assert sum(len(s.keep) for s in shards) == len(labels) and all(len(s.keep) == C * per // world for s in shards)
    np.random.seed(11)
    real_idx = shards[0].sample_all_classes(n)
    np.random.seed(11)
    assert np.array_equal(real_idx, shards[3].sample_all_classes(n))          # every rank replays the same stream
    seen = []
    sums = torch.zeros(C)
    for s in shards:
        loc, offs = s.local_sample(real_idx)
        offs = offs.tolist()
        assert offs[0] == 0 and offs[-1] == loc.numel() and len(offs) == C + 1
        vals = s.videos[loc][:, 0, 0, 0, 0]                  # = global index of the row
        for c in range(C):
            seg = vals[offs[c]:offs[c + 1]].long().tolist()
            assert all(labels[g] == c for g in seg)
            assert seg == [g for g in real_idx[c].tolist() if g in set(seg)]   # sampled order kept
            seen += seg
            sums[c] += float(sum(seg))
            assert abs(len(seg) - n / world) <= 3
    assert sorted(seen) == sorted(real_idx.reshape(-1).tolist())
    assert torch.equal(sums, torch.tensor([float(real_idx[c].sum()) for c in range(C)]))
    # gather plan: padded per-rank blocks, concatenated in rank order, indexed back into sample order
    n_max, gidx = shards[1].gather_plan(real_idx)
    blocks = []
    for s in shards:
        loc, _ = s.local_sample(real_idx)
        v = s.videos[loc][:, 0, 0, 0, 0]
        assert v.numel() <= n_max
        blocks.append(torch.cat([v, torch.full((n_max - v.numel(),), -1.0)]))
    gathered = torch.cat(blocks)
    assert torch.equal(gathered[gidx].long().view(C, n), torch.from_numpy(real_idx))
    assert all(torch.equal(s.gather_plan(real_idx)[1], gidx) for s in shards)
