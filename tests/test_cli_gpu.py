"""End-to-end runs of the drop-in driver scripts (reference flags) on a tiny synthetic dataset: buffer.py ->
replay_buffer_0.pt -> distill_s2d_ms.py --method MTT / DM and distill_baseline.py --method DM / MTT."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DATA = 'synthetic-3x6x8x64'          # 3 classes x 6 videos of 8x3x64x64


def _run(parser, main, argv):
    from video_distillation_b200 import cli
    args = getattr(cli, parser)().parse_args(argv.split())
    return getattr(cli, main)(args)


def test_buffer_then_mtt_and_dm(tmp_path):
    buf, save = str(tmp_path / 'buffers'), str(tmp_path / 'out')
    torch.manual_seed(0)
    np.random.seed(0)
    _run('buffer_parser', 'main_buffer', f'--dataset {DATA} --num_experts 1 --train_epochs 2 --save_interval 1 --batch_train 8 '
                                          f'--lr_teacher 0.01 --buffer_path {buf} --precision fp32')
    traj = torch.load(os.path.join(buf, 'replay_buffer_0.pt'), weights_only=False)
    assert len(traj) == 1 and len(traj[0]) == 3 and len(traj[0][0]) == 8                 # 1 expert, 2 epochs + init, 8 tensors
    assert tuple(traj[0][0][0].shape) == (64, 3, 3, 7, 7) and tuple(traj[0][0][6].shape) == (3, 128, 1, 1, 1)
    assert all(not t.is_cuda for t in traj[0][1]) and not torch.equal(traj[0][0][2], traj[0][2][2])

    # MTT + S2D for two iterations (no evaluation: startIt beyond Iteration), exact fp32 kernels
    tr = _run('s2d_parser', 'main_s2d', f'--method MTT --dataset {DATA} --vpc 1 --spc 2 --dpc 2 --frames 8 --syn_steps 2 --expert_epochs 1 '
                                        f'--max_start_epoch 1 --lr_teacher 0.01 --lr_dynamic 10 --lr_hal 0.01 --no_train_static --train_lr '
                                        f'--Iteration 1 --startIt 5 --buffer_path {buf} --save_path {save} --precision fp32')
    assert torch.isfinite(tr.dynamic_syn).all() and tr.dynamic_syn.grad.abs().sum() > 0 and float(tr.syn_lr.detach()) >= 0.001

    # DM + S2D on the tensor-core path with one (1-epoch) evaluation and the reference's checkpoint files
    tr = _run('s2d_parser', 'main_s2d', f'--method DM --dataset {DATA} --vpc 1 --spc 2 --dpc 2 --frames 8 --batch_real 4 --no_train_static '
                                        f'--lr_dynamic 10 --lr_hal 0.01 --Iteration 2 --eval_it 100 --num_eval 1 --epoch_eval_train 1 '
                                        f'--batch_train 4 --save_path {save} --run_name t')
    d = os.path.join(save, 'S2D_multis_DM', 't')
    dyn = torch.load(os.path.join(d, 'dynamic_0.pt'))
    hal = torch.load(os.path.join(d, 'hal_0.pt'))
    assert tuple(dyn.shape) == (6, 8, 1, 64, 64) and set(hal) == {'0.encoder.weight', '0.encoder.bias'}
    assert os.path.exists(os.path.join(d, 'weights_best.pt')) and not os.path.exists(os.path.join(d, 'images_0.pt'))
    assert tr.dynamic_syn.grad.abs().sum() > 0


def test_baseline_dm_and_mtt(tmp_path):
    buf, save = str(tmp_path / 'buffers'), str(tmp_path / 'out')
    _run('buffer_parser', 'main_buffer', f'--dataset {DATA} --num_experts 1 --train_epochs 1 --save_interval 1 --batch_train 8 '
                                          f'--lr_teacher 0.01 --buffer_path {buf}')
    np.random.seed(3)
    tr = _run('baseline_parser', 'main_baseline', f'--method DM --dataset {DATA} --ipc 1 --frames 8 --batch_real 4 --init real --lr_img 1 '
                                                   f'--Iteration 1 --eval_it 100 --num_eval 1 --epoch_eval_train 1 --batch_train 3 '
                                                   f'--save_path {save} --run_name b')
    img = torch.load(os.path.join(save, 'Baseline_DM', 'b', 'images_0.pt'))
    assert tuple(img.shape) == (3, 8, 3, 64, 64) and tr.image_syn.grad.abs().sum() > 0
    tr = _run('baseline_parser', 'main_baseline', f'--method MTT --dataset {DATA} --ipc 1 --frames 8 --init noise --syn_steps 2 --expert_epochs 1 '
                                                   f'--max_start_epoch 1 --lr_teacher 0.01 --lr_img 10 --Iteration 0 --eval_it 100 --num_eval 1 '
                                                   f'--epoch_eval_train 1 --batch_train 3 --buffer_path {buf} --save_path {save} --run_name m '
                                                   f'--precision fp32')
    assert torch.isfinite(tr.image_syn).all() and tr.image_syn.grad.abs().sum() > 0


def test_coreset_driver(tmp_path):
    img, lab, chosen, best = _run('coreset_parser', 'main_coreset', f'--dataset {DATA} --method herding --ipc 2 --frames 8 --num_eval 1 '
                                                                    f'--epoch_eval_train 1 --batch_train 6 --lr_net 0.01')
    assert tuple(img.shape) == (6, 8, 3, 64, 64) and lab.tolist() == [0, 0, 1, 1, 2, 2]
    assert all(i // 6 == c for i, c in zip(chosen, lab.tolist())) and len(set(chosen)) == 6
