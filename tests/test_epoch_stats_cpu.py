"""EpochStats (device-side running statistics of utils.epoch) against a numpy restatement of the per-batch bookkeeping of
the reference's epoch() (utils.py:775-845)."""
from collections import defaultdict

import numpy as np
import pytest
import torch

from video_distillation_b200.utils import EpochStats


def reference_bookkeeping(batches, train):
    loss_avg = acc_avg = num_exp = 0
    top = {1: 0.0, 3: 0.0, 5: 0.0}
    per = defaultdict(list)
    for out, lab, loss in batches:
        o, l = out.numpy(), lab.numpy()
        matched = np.equal(np.argmax(o, axis=-1), l)
        order = np.argsort(o, axis=-1)
        for k in ((5,) if train else (1, 3, 5)):
            top[k] += float(np.sum([l[i] in order[i, -k:] for i in range(len(l))]))
        for y, c in zip(l.tolist(), matched.tolist()):
            per[y].append(c)
        loss_avg += loss.item() * len(l)
        acc_avg += float(np.sum(matched))
        num_exp += len(l)
    per = dict(per)
    per = [np.mean(per[i]) if i in per else None for i in range(len(per))]
    return loss_avg / num_exp, [acc_avg / num_exp, top[1] / num_exp, top[3] / num_exp, top[5] / num_exp], per


@pytest.mark.parametrize('C,train', [(50, False), (50, True), (3, False), (7, True)])
def test_epoch_stats_match_reference_bookkeeping(C, train):
    g = torch.Generator().manual_seed(C)
    batches = []
    for n in (16, 16, 5):
        out = torch.randn(n, C, generator=g)
        lab = torch.randint(0, C if C < 10 else 12, (n,), generator=g)        # C=50: only some classes are seen
        batches.append((out, lab, torch.rand((), generator=g)))
    st = EpochStats(C, 'cpu')
    for out, lab, loss in batches:
        st.add(out, lab, loss, train)
    loss, accs, per = st.result(True)
    rl, ra, rp = reference_bookkeeping(batches, train)
    assert abs(loss - rl) < 1e-6 and np.allclose(accs, ra, atol=1e-12)
    assert len(per) == len(rp) and all((a is None and b is None) or abs(a - b) < 1e-12 for a, b in zip(per, rp))
    loss2, acc2, per2 = st.result(False)
    assert acc2 == accs[0] and per2 == per


@pytest.mark.parametrize('tag,C,train', [('test50', 50, False), ('train50', 50, True), ('test3', 3, False), ('train7', 7, True)])
def test_epoch_stats_match_the_live_reference_golden(tag, C, train):
    """tests/golden/epoch_stats.npz holds the outputs of the reference's own utils.epoch (run by oracle/make_golden.py with a stub
    network returning hash-generated logits); EpochStats must reproduce them from the same logits."""
    import os
    import torch.nn.functional as F
    from oracle import synth
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'epoch_stats.npz'))
    batches = []
    for i, n in enumerate((16, 16, 5)):
        logits = synth.hash_uniform((n, C), 400 + C + 10 * i, 3.0)
        labels = (synth.hash_uniform((n,), 400 + C + 10 * i + 1, 0.5) + 0.5).mul(min(C, 12)).long().clamp_(0, C - 1)
        labels[::2] = logits[::2].argmax(-1)
        labels[1::4] = logits[1::4].topk(min(3, C), dim=-1).indices[:, -1]
        batches.append((logits, labels))
    st = EpochStats(C, 'cpu')
    for _ in range(1 if train else 3):                      # the test branch of the reference makes three passes
        for logits, labels in batches:
            st.add(logits, labels, F.cross_entropy(logits, labels), train)
    loss, accs, per = st.result(True)
    assert abs(loss - float(gold[tag + '_loss'])) < 1e-6
    assert np.allclose(accs, gold[tag + '_accs'], atol=1e-12)
    ref_per = gold[tag + '_per_class']
    assert len(per) == len(ref_per)
    for a, b in zip(per, ref_per):
        assert (a is None and np.isnan(b)) or abs(a - b) < 1e-12
