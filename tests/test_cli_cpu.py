"""The driver scripts keep the reference's argument surface (tests/golden/cli_flags.json, generated from the reference's
own argparse blocks by oracle/make_cli_golden.py) and the on-disk formats of the expert buffers."""
import json
import os

import pytest
import torch

from video_distillation_b200 import cli
from video_distillation_b200.datasets import get_dataset

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'cli_flags.json')))
EXTRA = {'--precision', '--run_name'}


@pytest.mark.parametrize('script,make', [('distill_s2d_ms.py', cli.s2d_parser), ('distill_baseline.py', cli.baseline_parser),
                                         ('buffer.py', cli.buffer_parser), ('distill_coreset.py', cli.coreset_parser)])
def test_flags_match_reference(script, make):
    parser = make()
    ours = {a.option_strings[0]: a for a in parser._actions if a.option_strings and a.dest != 'help'}
    ref = GOLDEN[script]
    assert set(ours) - EXTRA == set(ref), (sorted(set(ref) - set(ours)), sorted(set(ours) - EXTRA - set(ref)))
    for flag, spec in ref.items():
        a = ours[flag]
        assert a.default == spec['default'], (flag, a.default, spec['default'])
        assert getattr(a.type, '__name__', None) == spec['type'], flag
        assert type(a).__name__ == spec['action'], flag
        assert (list(a.choices) if a.choices else None) == spec['choices'], flag


def test_reference_command_lines_parse():
    # sh/s2d/s2d_DM_ms.sh and sh/baseline/DM.sh style invocations
    a = cli.s2d_parser().parse_args('--method DM --dataset miniUCF101 --vpc 1 --spc 2 --dpc 2 --batch_real 64 --no_train_static '
                                    '--lr_dynamic 1e4 --lr_hal 1e-2 --eval_it 500 --preload --frames 16'.split())
    assert a.method == 'DM' and a.no_train_static and a.lr_dynamic == 1e4 and a.precision == 'f16x3r2'
    b = cli.baseline_parser().parse_args('--method DM --ipc 1 --batch_real 64 --init real --model ConvNet3D --frames 16'.split())
    assert b.ipc == 1 and b.init == 'real' and b.lr_img == 1


def test_synthetic_dataset_tuple(tmp_path):
    channel, im_size, C, names, mean, std, dst_train, dst_test, testloader = get_dataset('synthetic-3x4x4x16', str(tmp_path))
    assert (channel, im_size, C) == (3, (16, 16), 3) and len(dst_train) == 12 and len(names) == 3
    x, y = dst_train[5]
    assert x.shape == (4, 3, 16, 16) and int(y) == 1 and dst_train.labels[5] == 1
    # tensor-file form
    torch.save({'images_train': dst_train.images, 'labels_train': torch.tensor(dst_train.labels), 'images_test': dst_test.images,
                'labels_test': torch.tensor(dst_test.labels)}, tmp_path / 'mine.pt')
    t = get_dataset('mine', str(tmp_path))
    assert t[2] == 3 and torch.equal(t[6].images, dst_train.images)
    with pytest.raises(NotImplementedError):
        get_dataset('HMDB51', str(tmp_path))


def test_expert_buffers_walk(tmp_path):
    # replay_buffer_{n}.pt: list[trajectory] of list[epoch] of list[tensor] (buffer.py:75-103)
    import random
    import numpy as np
    for n in range(2):
        traj = [[[torch.full((2,), float(100 * n + 10 * t + e)) for _ in range(3)] for e in range(4)] for t in range(2)]
        torch.save(traj, tmp_path / f'replay_buffer_{n}.pt')
    random.seed(0)
    np.random.seed(0)
    eb = cli.ExpertBuffers(str(tmp_path), max_start_epoch=2, expert_epochs=1)
    seen = []
    for _ in range(5):
        start, target, e0 = eb.draw()
        assert 0 <= e0 < 2 and float(target[0][0]) - float(start[0][0]) == 1.0
        seen.append(float(start[0][0]))
    assert len(set(int(v) // 100 for v in seen)) == 1          # like the reference, the resident buffer is only reshuffled


def test_expert_buffers_reproduce_the_reference_walk(tmp_path):
    """tests/golden/expert_walk.npz: (start tensor, target tensor, start epoch) of nine consecutive iterations drawn by the
    reference's own buffer statements (executed verbatim by oracle/make_golden.py) with random.seed(11), np.random.seed(12)."""
    import random
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'expert_walk.npz'))['walk']
    for n in range(2):
        traj = [[[torch.full((2,), float(100 * n + 10 * t + e)) for _ in range(3)] for e in range(5)] for t in range(3)]
        torch.save(traj, tmp_path / f'replay_buffer_{n}.pt')
    random.seed(11)
    np.random.seed(12)
    eb = cli.ExpertBuffers(str(tmp_path), max_start_epoch=3, expert_epochs=1)
    walk = []
    for _ in range(9):
        start, target, e0 = eb.draw()
        walk.append((float(start[0][0]), float(target[0][0]), e0))
    assert np.array_equal(np.asarray(walk, dtype=np.float64), gold)
