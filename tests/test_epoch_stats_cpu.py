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
