"""Command-line drivers with the UNCHANGED argument surface of the reference scripts

    distill_s2d_ms.py   (:451-506)   DM / MTT with static + dynamic memory
    distill_baseline.py (:366-417)   DM / MTT on leaf synthetic videos
    buffer.py           (:107-131)   expert-trajectory producer (replay_buffer_{n}.pt)

Same flags, defaults, evaluation / checkpoint cadence and file formats (``images_{it}.pt``, ``dynamic_{it}.pt`` =
``dynamic_syn.flatten(0, 1)``, ``hal_{it}.pt`` = state_dict of the hallucinator ModuleList, ``weights_best.pt``, ``replay_buffer_{n}.pt`` =
``list[trajectory] of list[epoch] of list[8 CPU tensors]``), but the loop bodies are the batched trainers of
``distill.py`` on the hand-written kernels, one process per GPU (launch with torchrun to shard classes across GPUs;
rank 0 evaluates and saves).  Not reproduced: wandb logging (plain prints; ``wandb.run.name`` in the save path becomes
``--run_name``), the DC method and DSA / ZCA options (off the hot path).  Two optional flags are added:
``--precision`` (f16x3 parity mode of the tensor-core path | bf16 throughput mode | fp32 exact CUDA-core path) and ``--run_name``.

Multi-process runs (torchrun): ``init_distributed`` creates the process group with a long timeout; all ranks share one base
seed (rank 0's, broadcast), the initial memories are broadcast from rank 0, every iteration re-seeds all RNG streams from
(base seed, iteration) — rank 0 alone consumes random numbers while it evaluates — and the ranks meet at a barrier after each
evaluation block, so the replicas of the synthetic memories stay identical.
"""
import argparse
import copy
import datetime
import os
import random

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .distill import (DeviceDataset, DMBaselineTrainer, DMS2DTrainer, MTTBaselineTrainer, MTTS2DTrainer, _world)
from .utils import (Conv3DNet, MultiStaticSharedDataset, epoch, evaluate_synset, get_dataset, get_eval_pool, get_loops,
                    get_network, get_time)


# ------------------------------------------------------------------------------------------ parsers
def _extra(parser):
    parser.add_argument('--precision', type=str, default='f16x3r2', choices=['f16x3r2', 'f16x3', 'bf16', 'fp32', 'bf16x3'],
                        help='[B200] f16x3r2 (default): fused tensor-core pipeline on fp16 hi/lo operand pairs, the parity mode — three products '
                             'per MAC for the differentiable synthetic videos, two (exact weights, fp16 activations) for the frozen real videos '
                             'of the DM loss; f16x3: three products everywhere; bf16: single-pass throughput mode; fp32: exact CUDA-core '
                             'kernels; bf16x3: unfused conv trio with every fprop / dgrad / wgrad on bf16 hi/lo pairs.  The MTT unroll runs the '
                             'bf16x3 trio (its parity mode) unless bf16 (split fprop, single-pass dgrad / wgrad) or fp32 is given')
    parser.add_argument('--run_name', type=str, default=None, help='[B200] directory name under save_path/<project> (wandb.run.name in the reference)')
    return parser


def s2d_parser():
    """distill_s2d_ms.py:451-503, flag for flag."""
    parser = argparse.ArgumentParser(description='Parameter Processing')
    parser.add_argument('--dataset', type=str, default='miniUCF101', help='dataset')
    parser.add_argument('--method', type=str, default='MTT', help='MTT or DC or DM')
    parser.add_argument('--model', type=str, default='ConvNet3D', help='model')
    parser.add_argument('--spc', type=int, default=10, help='static memory(s) per class')
    parser.add_argument('--dpc', type=int, default=1, help='dynamic memory(s) per class')
    parser.add_argument('--vpc', type=int, default=5, help='')
    parser.add_argument('--eval_mode', type=str, default='S', help='eval_mode, check utils.py for more info')
    parser.add_argument('--num_eval', type=int, default=5, help='how many networks to evaluate on')
    parser.add_argument('--eval_it', type=int, default=100, help='how often to evaluate')
    parser.add_argument('--epoch_eval_train', type=int, default=1000, help='epochs to train a model with synthetic data')
    parser.add_argument('--Iteration', type=int, default=15000, help='how many distillation steps to perform')
    parser.add_argument('--no_train_static', action='store_true', help='do not train static memory')
    parser.add_argument('--path_static', type=str, default=None, help='path to pretrained static memory')
    parser.add_argument('--lr_static', type=float, default=100, help='learning rate for updating synthetic static memory')
    parser.add_argument('--lr_dynamic', type=float, default=0.01, help='learning rate for updating synthetic dynamic memory')
    parser.add_argument('--train_lr', action='store_true', help='train the learning rate')
    parser.add_argument('--lr_lr', type=float, default=1e-05, help='learning rate for updating... learning rate')
    parser.add_argument('--lr_teacher', type=float, default=0.01, help='initialization for synthetic learning rate')
    parser.add_argument('--lr_hal', type=float, default=0.01, help='learning rate for updating hallucinator')
    parser.add_argument('--batch_real', type=int, default=256, help='batch size for real data')
    parser.add_argument('--batch_syn', type=int, default=None, help='should only use this if you run out of VRAM')
    parser.add_argument('--batch_train', type=int, default=256, help='batch size for training networks')
    parser.add_argument('--data_path', type=str, default='distill_utils/data', help='dataset path')
    parser.add_argument('--buffer_path', type=str, default='./buffers', help='buffer path')
    parser.add_argument('--expert_epochs', type=int, default=3, help='how many expert epochs the target params are')
    parser.add_argument('--syn_steps', type=int, default=64, help='how many steps to take on synthetic data')
    parser.add_argument('--max_start_epoch', type=int, default=25, help='max epoch we can start at')
    parser.add_argument('--preload', action='store_true', help='preload all data into RAM')
    parser.add_argument('--n_hal', type=int, default=1, help='number of hallucinators')
    parser.add_argument('--frames', type=int, default=16, help='number of frames')
    parser.add_argument('--num_workers', type=int, default=8, help='number of workers')
    parser.add_argument('--startIt', type=int, default=0, help='start iteration')
    parser.add_argument('--save_path', type=str, default='./logged_files', help='path to result')
    return _extra(parser)


def baseline_parser():
    """distill_baseline.py:366-413, flag for flag."""
    parser = argparse.ArgumentParser(description='Parameter Processing')
    parser.add_argument('--dataset', type=str, default='miniUCF101', help='dataset')
    parser.add_argument('--method', type=str, default='DC', help='MTT or DM')
    parser.add_argument('--model', type=str, default='ConvNet3D', help='model')
    parser.add_argument('--ipc', type=int, default=1, help='image(s) per class')
    parser.add_argument('--eval_mode', type=str, default='S', help='use top5 to eval top5 accuracy, use S to eval single accuracy')
    parser.add_argument('--outer_loop', type=int, default=None, help='')
    parser.add_argument('--inner_loop', type=int, default=None, help='')
    parser.add_argument('--num_eval', type=int, default=5, help='how many networks to evaluate on')
    parser.add_argument('--eval_it', type=int, default=50, help='how often to evaluate')
    parser.add_argument('--epoch_eval_train', type=int, default=1000, help='epochs to train a model with synthetic data')
    parser.add_argument('--Iteration', type=int, default=1000, help='how many distillation steps to perform')
    parser.add_argument('--lr_net', type=float, default=0.001, help='learning rate for network')
    parser.add_argument('--lr_img', type=float, default=1, help='learning rate for synthetic data')
    parser.add_argument('--lr_lr', type=float, default=1e-5, help='learning rate for synthetic data')
    parser.add_argument('--lr_teacher', type=float, default=0.001, help='learning rate for teacher')
    parser.add_argument('--train_lr', action='store_true', help='train synthetic lr')
    parser.add_argument('--batch_real', type=int, default=256, help='batch size for real data')
    parser.add_argument('--batch_train', type=int, default=256, help='batch size for training networks')
    parser.add_argument('--batch_syn', type=int, default=None, help='batch size for syn')
    parser.add_argument('--init', type=str, default='real', choices=['noise', 'real', 'real-all'],
                        help='noise/real: initialize synthetic images from random noise or randomly sampled real images.')
    parser.add_argument('--data_path', type=str, default='distill_utils/data', help='dataset path')
    parser.add_argument('--expert_epochs', type=int, default=3, help='how many expert epochs the target params are')
    parser.add_argument('--syn_steps', type=int, default=64, help='how many steps to take on synthetic data')
    parser.add_argument('--max_start_epoch', type=int, default=25, help='max epoch we can start at')
    parser.add_argument('--dis_metric', type=str, default='ours', help='distance metric')
    parser.add_argument('--buffer_path', type=str, default=None, help='buffer path')
    parser.add_argument('--num_workers', type=int, default=8, help='')
    parser.add_argument('--preload', action='store_true', help='preload dataset')
    parser.add_argument('--save_path', type=str, default='./logged_files', help='path to save')
    parser.add_argument('--frames', type=int, default=16, help='')
    return _extra(parser)


def buffer_parser():
    """buffer.py:107-128, flag for flag."""
    parser = argparse.ArgumentParser(description='Parameter Processing')
    parser.add_argument('--dataset', type=str, default='miniUCF101', help='dataset')
    parser.add_argument('--model', type=str, default='ConvNet3D', help='model')
    parser.add_argument('--num_experts', type=int, default=100, help='training iterations')
    parser.add_argument('--lr_teacher', type=float, default=0.001, help='learning rate for updating network parameters')
    parser.add_argument('--batch_train', type=int, default=256, help='batch size for training networks')
    parser.add_argument('--batch_real', type=int, default=256, help='batch size for real loader')
    parser.add_argument('--num_workers', type=int, default=8, help='')
    parser.add_argument('--data_path', type=str, default='distill_utils/data', help='dataset path')
    parser.add_argument('--buffer_path', type=str, default='./logs/buffers', help='buffer path')
    parser.add_argument('--train_epochs', type=int, default=50)
    parser.add_argument('--decay', action='store_true')
    parser.add_argument('--mom', type=float, default=0, help='momentum')
    parser.add_argument('--l2', type=float, default=0, help='l2 regularization')
    parser.add_argument('--save_interval', type=int, default=10)
    parser.add_argument('--preload', action='store_true', help='preload dataset to memory')
    return _extra(parser)


def coreset_parser():
    """distill_coreset.py:148-167, flag for flag."""
    parser = argparse.ArgumentParser(description='Parameter Processing')
    parser.add_argument('--dataset', type=str, default='miniUCF101', help='dataset')
    parser.add_argument('--method', type=str, default='k-center', help='k-center or herding')
    parser.add_argument('--model', type=str, default='ConvNet3D', help='model')
    parser.add_argument('--ipc', type=int, default=1, help='image(s) per class')
    parser.add_argument('--eval_mode', type=str, default='S', help='eval_mode, check utils.py for more info')
    parser.add_argument('--num_eval', type=int, default=5, help='how many networks to evaluate on')
    parser.add_argument('--epoch_eval_train', type=int, default=1000, help='epochs to train a model with synthetic data')
    parser.add_argument('--lr_net', type=float, default=0.001, help='learning rate for network')
    parser.add_argument('--batch_train', type=int, default=256, help='batch size for training networks')
    parser.add_argument('--data_path', type=str, default='distill_utils/data', help='dataset path')
    parser.add_argument('--pretrained_path', type=str, default=None, help='pretrained model path')
    parser.add_argument('--num_workers', type=int, default=8, help='')
    parser.add_argument('--save_path', type=str, default='.', help='path to save')
    parser.add_argument('--frames', type=int, default=16, help='')
    parser.add_argument('--preload', action='store_true', help='preload dataset')
    return _extra(parser)


# ------------------------------------------------------------------------------------------ shared pieces
def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('video_distillation_b200 drivers need a CUDA device (there is no CPU fallback)')
    rank, world = _world()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    return torch.device('cuda', local), rank, world


def init_distributed():
    """torchrun entry: NCCL process group with a LONG timeout — rank 0 alone runs the evaluation blocks (num_eval networks x
    epoch_eval_train epochs) while the other ranks wait at the barrier behind them, far longer than NCCL's default 10 minutes."""
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        local = int(os.environ.get('LOCAL_RANK', '0'))
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = dict(device_id=torch.device('cuda', local)) if backend == 'nccl' else {}
        dist.init_process_group(backend, timeout=datetime.timedelta(hours=24), **kw)


def _seed_all(seed):
    seed = int(seed) % (2 ** 31 - 1)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _shared_seed(world, dev):
    """One base seed for ALL ranks (drawn on rank 0 from its entropy-seeded generator, broadcast): every rank must draw the
    same static / dynamic memories, hallucinator, init='real' videos, expert trajectories and per-iteration indices — the
    trainers only shard the WORK of an iteration, every rank applies the full update to its replica.  Single process: None
    (the RNG streams stay exactly the reference's)."""
    if world == 1:
        return None
    import torch.distributed as dist
    t = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)
    t = t.to(dev) if dist.get_backend() == 'nccl' else t
    dist.broadcast(t, src=0)
    return int(t.item())


def _broadcast_state(tensors, world):
    """Rank 0's initial values become everybody's (belt and braces on top of the shared seed)."""
    if world == 1:
        return
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t.data, src=0)


def _replicas_agree(tensors, world):
    """max |x_rank - x_0| over the given tensors and all ranks (diagnostic; 0.0 for a single process)."""
    if world == 1:
        return 0.0
    import torch.distributed as dist
    worst = 0.0
    for t in tensors:
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        d = (t.detach() - ref).abs().max().reshape(1).float()
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        worst = max(worst, float(d.item()))
    return worst


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _tensors_of(dst):
    """(videos, labels) of a dataset object: TensorDataset members, or one pass over __getitem__ (what --preload does,
    distill_s2d_ms.py:28-38)."""
    if hasattr(dst, 'images') and torch.is_tensor(dst.images):
        return dst.images, torch.as_tensor(dst.labels).long()
    xs, ys = [], []
    for i in range(len(dst)):
        x, y = dst[i]
        xs.append(x)
        ys.append(int(y))
    return torch.stack(xs), torch.tensor(ys)


def _mtt_precision(args):
    """Conv-trio precision of the MTT unroll: the parity-grade tensor-core trio (every primitive on bf16 hi / lo pairs) unless the
    single-pass throughput mode or the exact fp32 kernels are asked for."""
    return {'fp32': 'fp32', 'bf16': 'bf16'}.get(args.precision, 'bf16x3')


def _eval_precision(args):
    return 'bf16' if args.precision in ('bf16', 'bf16x3', 'f16x3', 'f16x3r2') else 'fp32'


class ExpertBuffers:
    """Expert-trajectory bookkeeping of the MTT loops (distill_s2d_ms.py:117-133, 203-229): files are discovered as
    replay_buffer_{n}.pt, shuffled with ``random``, and a (trajectory, start epoch) pair is drawn per iteration.  Like the
    reference, advancing to the next file reshuffles the resident buffer (the reference never reloads it, :219-223)."""

    def __init__(self, expert_dir, max_start_epoch, expert_epochs):
        self.files, n = [], 0
        while os.path.exists(os.path.join(expert_dir, 'replay_buffer_{}.pt'.format(n))):
            self.files.append(os.path.join(expert_dir, 'replay_buffer_{}.pt'.format(n)))
            n += 1
        if n == 0:
            raise AssertionError('No buffers detected at {}'.format(expert_dir))
        self.file_idx, self.expert_idx = 0, 0
        self.max_start_epoch, self.expert_epochs = max_start_epoch, expert_epochs
        random.shuffle(self.files)
        print('loading file {}'.format(self.files[self.file_idx]))
        self.buffer = torch.load(self.files[self.file_idx], weights_only=False)
        random.shuffle(self.buffer)

    def draw(self):
        traj = self.buffer[self.expert_idx]
        self.expert_idx += 1
        if self.expert_idx == len(self.buffer):
            self.expert_idx = 0
            self.file_idx += 1
            if self.file_idx == len(self.files):
                self.file_idx = 0
                random.shuffle(self.files)
            print('loading file {}'.format(self.files[self.file_idx]))
            random.shuffle(self.buffer)
        start_epoch = np.random.randint(0, self.max_start_epoch)
        return traj[start_epoch], traj[start_epoch + self.expert_epochs], start_epoch


def _evaluate(args, it, model_eval_pool, channel, num_classes, im_size, payload, testloader, mode, best_acc, best_std, test_freq=None):
    """The evaluation block shared by the four loops (distill_s2d_ms.py:137-173 / distill_baseline.py:274-303)."""
    save_this_it = False
    eargs = copy.copy(args)
    eargs.precision = _eval_precision(args)
    for model_eval in model_eval_pool:
        print('-------------------------\nEvaluation\nmodel_train = %s, model_eval = %s, iteration = %d' % (args.model, model_eval, it))
        accs_test, accs_train = [], []
        for it_eval in range(args.num_eval):
            net_eval = get_network(model_eval, channel, num_classes, im_size, frames=args.frames).to(args.device)
            images, labels = payload()
            _, acc_train, acc_test, acc_per_cls = evaluate_synset(it_eval, net_eval, images, labels, testloader, eargs, mode=mode,
                                                                  test_freq=test_freq)
            accs_test.append(acc_test)
            accs_train.append(acc_train)
        accs_test = np.array(accs_test)
        acc_test_mean, acc_test_std = np.mean(accs_test), np.std(accs_test)
        if acc_test_mean > best_acc[model_eval]:
            best_acc[model_eval], best_std[model_eval] = acc_test_mean, acc_test_std
            save_this_it = True
        print('Evaluate %d random %s, mean = %.4f std = %.4f\n-------------------------' % (len(accs_test), model_eval, acc_test_mean, acc_test_std))
    return save_this_it


# ------------------------------------------------------------------------------------------ distill_s2d_ms.py
def main_s2d(args):
    """distill_s2d_ms.py:18-448 on the batched trainers."""
    dev, rank, world = _require_cuda()
    args.device = str(dev)
    eval_it_pool = np.arange(args.startIt, args.Iteration + 1, args.eval_it).tolist()
    print('Evaluation iterations: ', eval_it_pool)
    channel, im_size, num_classes, class_names, mean, std, dst_train, dst_test, testloader = get_dataset(args.dataset, args.data_path)
    model_eval_pool = get_eval_pool(args.eval_mode, args.model, args.model)
    project_name = 'S2D_multis_{}'.format(args.method)          # distill_s2d_ms.py:47
    run_name = args.run_name or f'{args.dataset}_vpc{args.vpc}_{datetime.datetime.now().strftime("%Y%m%d%H%M%S")}'
    if args.batch_syn is None:
        args.batch_syn = num_classes * args.vpc
    args.distributed = world > 1
    print('Hyper-parameters: \n', args.__dict__)
    print('Evaluation model pool: ', model_eval_pool)
    if args.n_hal != 1:
        raise NotImplementedError('the reference only ever applies hals[0] (distill_s2d_ms.py:412); n_hal must be 1')
    base_seed = _shared_seed(world, dev)
    if base_seed is not None:
        _seed_all(base_seed)

    static_syn = torch.randn(size=(num_classes * args.spc, 3, im_size[0], im_size[1]), dtype=torch.float)
    dynamic_syn = torch.randn(size=(num_classes, args.dpc, args.frames, 1, im_size[0], im_size[1]), dtype=torch.float)
    hal = Conv3DNet()
    if args.path_static is not None:
        static_syn = torch.load(args.path_static)['image']
        print('load static memory from %s' % args.path_static)
        print('static_syn shape: ', static_syn.shape)
    common = dict(num_classes=num_classes, channel=channel, im_size=im_size, frames=args.frames, vpc=args.vpc, spc=args.spc,
                  dpc=args.dpc, lr_dynamic=args.lr_dynamic, lr_hal=args.lr_hal, lr_static=args.lr_static,
                  train_static=not args.no_train_static, static_syn=static_syn, dynamic_syn=dynamic_syn, hal=hal, device=dev)
    print('%s training begins' % get_time())
    best_acc = {m: 0 for m in model_eval_pool}
    best_std = {m: 0 for m in model_eval_pool}

    if args.method == 'MTT':
        experts = ExpertBuffers(args.buffer_path, args.max_start_epoch, args.expert_epochs)
        tr = MTTS2DTrainer(syn_steps=args.syn_steps, lr_lr=args.lr_lr, lr_teacher=args.lr_teacher, train_lr=args.train_lr,
                           batch_syn=args.batch_syn, precision=_mtt_precision(args), **common)
    elif args.method == 'DM':
        videos, labels = _tensors_of(dst_train)
        ds = DeviceDataset(videos, labels, num_classes, dev, rank, world, shard='video' if world > 1 else 'class')
        prec = args.precision
        tr = DMS2DTrainer(ds, batch_real=args.batch_real, precision=prec, max_batch=640 if prec in ('bf16', 'f16x3', 'f16x3r2') else 128, **common)
        if tr.embedder.tc is not None:
            ds.prepack(tr.embedder.tc, extra_slots=len(tr.owned) * tr.vpc)
    else:
        raise NotImplementedError('Method {} not implemented'.format(args.method))

    # DM shards the memories by class (every rank sliced the same seeded full-size init); MTT keeps full replicas
    _broadcast_state(list(tr.hal.parameters()) + ([tr.static_syn, tr.dynamic_syn, tr.syn_lr] if args.method == 'MTT' else []), world)
    full = {}

    def memories():
        return full['m'] if 'm' in full else (tr.static_syn.detach(), tr.dynamic_syn.detach())

    def payload():
        st, dy = memories()
        return [copy.deepcopy(st), copy.deepcopy(dy), nn.ModuleList([copy.deepcopy(tr.hal)])], None

    def hals_state():
        # the reference saves the state_dict of `hals = nn.ModuleList([Conv3DNet()])` (:96, :186): keys '0.encoder.weight/bias'
        return nn.ModuleList([tr.hal]).state_dict()

    for it in range(0, args.Iteration + 1):
        save_this_it = False
        if it in eval_it_pool and args.method == 'DM' and world > 1:
            full['m'] = tr.full_memories()            # collective: every rank contributes its class shard
        if it in eval_it_pool and rank == 0:
            args.lr_net = tr.syn_lr.detach() if args.method == 'MTT' else torch.tensor(args.lr_teacher)
            save_this_it = _evaluate(args, it, model_eval_pool, channel, num_classes, im_size, payload, testloader, 'multi-static',
                                     best_acc, best_std)
        if it in eval_it_pool and (save_this_it or it % 1000 == 0) and rank == 0:
            with torch.no_grad():
                save_dir = os.path.join(args.save_path, project_name, run_name)
                os.makedirs(save_dir, exist_ok=True)
                image_save, dynamic_save = memories()
                dynamic_save = dynamic_save.flatten(0, 1)
                if not args.no_train_static:
                    torch.save(image_save.cpu(), os.path.join(save_dir, 'images_{}.pt'.format(it)))
                torch.save(hals_state(), os.path.join(save_dir, 'hal_{}.pt'.format(it)))
                torch.save(dynamic_save.cpu(), os.path.join(save_dir, 'dynamic_{}.pt'.format(it)))
                if save_this_it:
                    if not args.no_train_static:
                        torch.save(image_save.cpu(), os.path.join(save_dir, 'images_best.pt'))
                    torch.save(hals_state(), os.path.join(save_dir, 'weights_best.pt'))
                    torch.save(dynamic_save.cpu(), os.path.join(save_dir, 'dynamic_best.pt'))
        if base_seed is not None:
            if it in eval_it_pool:
                _barrier(world)                       # the other ranks wait here while rank 0 evaluates / saves
            _seed_all(base_seed + 7919 * (it + 1))    # rank 0's evaluation consumed RNG: realign every stream, every iteration
        if args.method == 'MTT':
            start, target, start_epoch = experts.draw()
            grand_loss = tr.step(start, target, net_seed=None if world == 1 else 1000003 + it)
            if it % 10 == 0:
                print('%s iter = %04d, param_loss = %.4f, param_dist = %.4f, grand_loss = %.4f (start epoch %d)' % (
                    get_time(), it, tr.last['param_loss'].item(), tr.last['param_dist'].item(), grand_loss.item(), start_epoch))
        else:
            loss = tr.step(net_seed=None if world == 1 else 1000003 + it)      # ranks must agree on the fresh frozen net
            if it % 10 == 0:
                print('%s iter = %04d, loss = %.4f' % (get_time(), it, loss.item() / num_classes))
    return tr


# ------------------------------------------------------------------------------------------ distill_baseline.py
def main_baseline(args):
    """distill_baseline.py:20-362 on the batched trainers."""
    if args.outer_loop is None and args.inner_loop is None:
        args.outer_loop, args.inner_loop = get_loops(args.ipc)
    elif args.outer_loop is None or args.inner_loop is None:
        raise ValueError(f'Please set neither or both outer/inner_loop: {args.outer_loop}, {args.inner_loop}')
    print('outer_loop = %d, inner_loop = %d' % (args.outer_loop, args.inner_loop))
    dev, rank, world = _require_cuda()
    args.device = str(dev)
    eval_it_pool = np.arange(0, args.Iteration + 1, args.eval_it).tolist()
    print('Evaluation iterations: ', eval_it_pool)
    channel, im_size, num_classes, class_names, mean, std, dst_train, dst_test, testloader = get_dataset(args.dataset, args.data_path)
    model_eval_pool = get_eval_pool(args.eval_mode, args.model, args.model)
    project_name = 'Baseline_{}'.format(args.method)
    run_name = args.run_name or f'{args.dataset}_ipc{args.ipc}_{args.lr_img}_{datetime.datetime.now().strftime("%Y%m%d%H%M%S")}'
    if args.batch_syn is None:
        args.batch_syn = num_classes * args.ipc
    args.distributed = world > 1
    print('Hyper-parameters: \n', args.__dict__)
    base_seed = _shared_seed(world, dev)
    if base_seed is not None:
        _seed_all(base_seed)
    videos, labels = _tensors_of(dst_train)
    ds = DeviceDataset(videos, labels, num_classes, dev, rank, world)
    image_syn = torch.randn(size=(num_classes * args.ipc, args.frames, channel, im_size[0], im_size[1]), dtype=torch.float)
    if args.init == 'real':
        print('initialize synthetic data from random real images')
        full = ds if world == 1 else DeviceDataset(videos, labels, num_classes, 'cpu')      # every class, same numpy stream
        for c in range(num_classes):                                                        # distill_baseline.py:96-100
            image_syn[c * args.ipc:(c + 1) * args.ipc] = full.get_images(c, args.ipc).detach().cpu()
    else:
        print('initialize synthetic data from random noise')
    print('%s training begins' % get_time())
    best_acc = {m: 0 for m in model_eval_pool}
    best_std = {m: 0 for m in model_eval_pool}
    if args.method == 'MTT':
        experts = ExpertBuffers(args.buffer_path, args.max_start_epoch, args.expert_epochs)
        tr = MTTBaselineTrainer(num_classes=num_classes, channel=channel, im_size=im_size, frames=args.frames, ipc=args.ipc,
                                syn_steps=args.syn_steps, lr_img=args.lr_img, lr_lr=args.lr_lr, lr_teacher=args.lr_teacher,
                                train_lr=args.train_lr, batch_syn=args.batch_syn, image_syn=image_syn, device=dev,
                                precision=_mtt_precision(args))
    elif args.method == 'DM':
        prec = args.precision
        tr = DMBaselineTrainer(ds, num_classes=num_classes, channel=channel, im_size=im_size, frames=args.frames, ipc=args.ipc,
                               batch_real=args.batch_real, lr_img=args.lr_img, precision=prec, image_syn=image_syn,
                               max_batch=640 if prec in ('bf16', 'f16x3', 'f16x3r2') else 128, device=dev)
    else:
        raise NotImplementedError('Method {} not implemented (DC is outside the B200 hot path)'.format(args.method))
    label_syn = torch.tensor(np.stack([np.ones(args.ipc) * i for i in range(0, num_classes)]), dtype=torch.long,
                             requires_grad=False, device=dev).view(-1)

    _broadcast_state([tr.image_syn] + ([tr.syn_lr] if args.method == 'MTT' else []), world)

    def payload():
        return tr.image_syn.detach().clone(), label_syn.detach().clone()

    for it in range(0, args.Iteration + 1):
        save_this_it = False
        if it in eval_it_pool and rank == 0:
            if args.method == 'MTT':
                args.lr_net = tr.syn_lr.detach()
            save_this_it = _evaluate(args, it, model_eval_pool, channel, num_classes, im_size, payload, testloader, 'none',
                                     best_acc, best_std, test_freq=100 if args.method == 'DM' else 200)     # distill_baseline.py:304 / :158
        if it in eval_it_pool and (save_this_it or it % 1000 == 0) and rank == 0:
            save_dir = os.path.join(args.save_path, project_name, run_name)
            os.makedirs(save_dir, exist_ok=True)
            image_save = tr.image_syn.detach()
            torch.save(image_save.cpu(), os.path.join(save_dir, 'images_{}.pt'.format(it)))
            if save_this_it:
                torch.save(image_save.cpu(), os.path.join(save_dir, 'images_best.pt'))
        if base_seed is not None:
            if it in eval_it_pool:
                _barrier(world)
            _seed_all(base_seed + 7919 * (it + 1))
        if args.method == 'MTT':
            start, target, start_epoch = experts.draw()
            grand_loss = tr.step(start, target, net_seed=None if world == 1 else 1000003 + it)
            if it % 10 == 0:
                print('%s iter = %04d, grand_loss = %.4f (start epoch %d)' % (get_time(), it, grand_loss.item(), start_epoch))
        else:
            loss = tr.step(net_seed=None if world == 1 else 1000003 + it)
            if it % 10 == 0:
                print('%s iter = %04d, loss = %.4f' % (get_time(), it, loss.item() / num_classes))
    return tr


# ------------------------------------------------------------------------------------------ buffer.py
def main_buffer(args):
    """buffer.py:14-104: train ``num_experts`` teachers, keep the parameters after every epoch, save every
    ``save_interval`` trajectories as replay_buffer_{n}.pt (list[traj] of list[epoch] of list[CPU tensors])."""
    dev, rank, world = _require_cuda()
    args.device = str(dev)
    channel, im_size, num_classes, class_names, mean, std, dst_train, dst_test, testloader = get_dataset(
        args.dataset, args.data_path)
    save_dir = args.buffer_path
    os.makedirs(save_dir, exist_ok=True)
    criterion = nn.CrossEntropyLoss().to(args.device)
    trajectories = []
    trainloader = torch.utils.data.DataLoader(dst_train, batch_size=args.batch_train, shuffle=True, num_workers=0)
    frames = int(dst_train[0][0].shape[0])
    prev = ops.set_conv_backend('tc' if _eval_precision(args) == 'bf16' else 'fp32')
    try:
        for it in range(0, args.num_experts):
            teacher_net = get_network(args.model, channel, num_classes, im_size, frames=frames).to(args.device)
            teacher_net.train()
            lr = args.lr_teacher
            teacher_optim = torch.optim.SGD(teacher_net.parameters(), lr=lr, momentum=args.mom, weight_decay=args.l2)
            teacher_optim.zero_grad()
            timestamps = [[p.detach().cpu() for p in teacher_net.parameters()]]
            lr_schedule = [args.train_epochs // 2 + 1]
            for e in range(args.train_epochs):
                train_loss, train_acc, _ = epoch('train', dataloader=trainloader, net=teacher_net, optimizer=teacher_optim,
                                                 criterion=criterion, args=args)
                if e % 10 == 0 or e == args.train_epochs - 1:
                    with torch.no_grad():
                        test_loss, test_acc, acc_per_cls = epoch('test', dataloader=testloader, net=teacher_net, optimizer=None,
                                                                 criterion=criterion, args=args)
                    print('Itr: {}\tEpoch: {}\tTrain Acc: {}\tTest Acc: {}'.format(it, e, train_acc, test_acc))
                timestamps.append([p.detach().cpu() for p in teacher_net.parameters()])
                if e in lr_schedule and args.decay:
                    lr *= 0.1
                    teacher_optim = torch.optim.SGD(teacher_net.parameters(), lr=lr, momentum=args.mom, weight_decay=args.l2)
                    teacher_optim.zero_grad()
            trajectories.append(timestamps)
            if len(trajectories) == args.save_interval:
                n = 0
                while os.path.exists(os.path.join(save_dir, 'replay_buffer_{}.pt'.format(n))):
                    n += 1
                print('Saving {}'.format(os.path.join(save_dir, 'replay_buffer_{}.pt'.format(n))))
                torch.save(trajectories, os.path.join(save_dir, 'replay_buffer_{}.pt'.format(n)))
                trajectories = []
    finally:
        ops.set_conv_backend(prev)


# ------------------------------------------------------------------------------------------ distill_coreset.py
def main_coreset(args):
    """distill_coreset.py:24-144: k-center / herding over ConvNet3D embeddings of every training video, then the usual
    evaluation of the selected set.  Embeddings: tensor-core path (precision bf16) or the exact fp32 kernels."""
    from .coreset import select_coreset
    from .tc import TcConvNet3D, tc_supported
    dev, rank, world = _require_cuda()
    args.device = str(dev)
    model_eval_pool = get_eval_pool(args.eval_mode, args.model, args.model)
    channel, im_size, num_classes, class_names, mean, std, dst_train, dst_test, testloader = get_dataset(args.dataset, args.data_path)
    videos, labels = _tensors_of(dst_train)
    videos = videos.to(dev)
    net = get_network(args.model, channel, num_classes, im_size, frames=args.frames).to(dev)
    for param in net.parameters():
        param.requires_grad = False
    if args.pretrained_path is not None:
        print('Loading pretrained model')
        net.load_state_dict(torch.load(args.pretrained_path))
    net.eval()
    if args.precision in ('bf16', 'f16x3', 'f16x3r2') and tc_supported(args.frames, im_size[0], im_size[1]):
        # coreset selection compares individual embeddings: three products per MAC ('f16x3r2' only shortens class MEANS)
        tc = TcConvNet3D(args.frames, im_size[0], im_size[1], dev, max_batch=256, split=args.precision.startswith('f16x3'))
        f = net.features
        tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
        embed = tc.embed
    else:
        embed = net.embed
    image_syn, label_syn, chosen = select_coreset(embed, videos, labels, num_classes, args.ipc, args.method)
    print('Synthetic data generated by %s: dataset indices %s' % (args.method, chosen))
    best_acc = {m: 0 for m in model_eval_pool}
    best_std = {m: 0 for m in model_eval_pool}
    _evaluate(args, 0, model_eval_pool, channel, num_classes, im_size, lambda: (image_syn.detach().clone(), label_syn.detach().clone()),
              testloader, 'none', best_acc, best_std, test_freq=100)
    return image_syn, label_syn, chosen, best_acc
