"""Coreset selection (vectorised, device-agnostic) against the loop restatement of distill_coreset.py in oracle/."""
import pytest
import torch

from oracle import coreset as ref
from video_distillation_b200.coreset import herding_select, k_center_select, select_coreset


@pytest.mark.parametrize('n,d,ipc', [(7, 5, 1), (20, 16, 5), (64, 32, 10), (5, 3, 5)])
def test_selection_matches_reference_loops(n, d, ipc):
    g = torch.Generator().manual_seed(n * 100 + ipc)
    f = torch.randn(n, d, generator=g)
    assert k_center_select(f, ipc) == ref.k_center(f, ipc)
    assert herding_select(f, ipc) == ref.herding(f, ipc)


def test_ties_and_duplicates():
    f = torch.tensor([[0., 0.], [1., 0.], [1., 0.], [-1., 0.], [0., 0.]])
    assert k_center_select(f, 3) == ref.k_center(f, 3)
    assert herding_select(f, 4) == ref.herding(f, 4)


def test_select_coreset_shapes():
    videos = torch.randn(12, 2, 3, 4, 4)
    labels = [0, 1, 2] * 4
    img, lab, idx = select_coreset(lambda v: v.flatten(1), videos, labels, 3, 2, 'herding')
    assert img.shape == (6, 2, 3, 4, 4) and lab.tolist() == [0, 0, 1, 1, 2, 2] and all(labels[i] == c for i, c in zip(idx, lab.tolist()))


@pytest.mark.parametrize('tag,n,d,ipc', [('a', 7, 5, 1), ('b', 20, 16, 5), ('c', 64, 32, 10), ('d', 5, 3, 5)])
def test_selection_matches_live_reference_golden(tag, n, d, ipc):
    """tests/golden/coreset.npz: indices chosen by the reference's own selection statements (executed verbatim by
    oracle/make_golden.py) on hash-generated feature matrices."""
    import os
    import numpy as np
    from oracle import synth
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'coreset.npz'))
    f = synth.hash_uniform((n, d), 700 + n)
    assert k_center_select(f, 1) == gold[f'{tag}_k-center'].tolist()       # the reference's k-center is only well defined for its first centre
    assert herding_select(f, ipc) == gold[f'{tag}_herding'].tolist()
