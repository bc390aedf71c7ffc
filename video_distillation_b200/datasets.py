"""``get_dataset`` for the distillation drivers (reference: utils.py:118-460, distill_utils/dataset.py).

The reference's JPEG / frame-folder loaders (HMDB51, UCF101, Kinetics400, SSv2 ...) are outside the hot path
(SURVEY.md §2, §8f rank 2).  The drivers need the 9-tuple ``(channel, im_size, num_classes, class_names, mean,
std, dst_train, dst_test, testloader)`` — this module returns it for

* ``synthetic[-<C>x<N>x<T>x<H>]`` / ``<name>-synthetic`` : seeded N(0,1) videos of the named dataset's shape
  (miniUCF101: 50 classes 16x3x112x112; Kinetics400: 400 classes 8x3x64x64), N training videos per class;
* a tensor file ``<data_path>/<dataset>.pt`` holding ``{'images_train' (N,T,3,H,W) float, 'labels_train' (N,),
  'images_test', 'labels_test', optional 'class_names', 'mean', 'std'}`` — what ``--preload`` builds in the
  reference (distill_s2d_ms.py:28-38), saved once by the user's own decoder.
Anything else raises with a pointer to this contract instead of silently decoding on the CPU.
"""
import os

import torch
from torch.utils.data import DataLoader

from .utils import TensorDataset

SHAPES = {            # name -> (classes, frames, im_size, train videos per class, test videos per class)
    'miniUCF101': (50, 16, 112, 96, 16),
    'UCF101': (101, 16, 112, 96, 16),
    'HMDB51': (51, 16, 112, 70, 30),
    'Kinetics400': (400, 8, 64, 64, 8),
    'SSv2': (174, 8, 64, 64, 8),
}


class _LabelledTensorDataset(TensorDataset):
    """TensorDataset with the ``.labels`` attribute the drivers read when --preload is off (distill_s2d_ms.py:72)."""

    def __init__(self, images, labels):
        super().__init__(images, labels)
        self.labels = [int(v) for v in labels]


def _synthetic(C, n_train, n_test, T, H, seed=0):
    g = torch.Generator().manual_seed(seed)
    xtr = torch.randn(C * n_train, T, 3, H, H, generator=g)
    ytr = torch.arange(C).repeat_interleave(n_train)
    xte = torch.randn(C * n_test, T, 3, H, H, generator=g)
    yte = torch.arange(C).repeat_interleave(n_test)
    return xtr, ytr, xte, yte


def get_dataset(dataset, data_path, num_workers=0, img_size=(112, 112), split_num=1, split_id=0, split_mode='mean', *,
                test_batch_size=64):
    """Reference argument order (utils.py:21).  ``img_size`` / ``split_*`` are accepted for call compatibility (the tensor-file
    and synthetic sources have their own shapes and are not split).  The test loader uses the reference's hard-wired batch
    size 64 (utils.py:459) — `epoch` normalises per batch, so test loss / accuracy depend on it; ``test_batch_size`` is a
    keyword-only override."""
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]        # utils.py:214-230 (ImageNet statistics)
    class_names = None
    path = os.path.join(data_path or '.', f'{dataset}.pt')
    if os.path.exists(path):
        blob = torch.load(path, map_location='cpu')
        xtr, ytr = blob['images_train'].float(), torch.as_tensor(blob['labels_train']).long()
        xte, yte = blob['images_test'].float(), torch.as_tensor(blob['labels_test']).long()
        class_names = blob.get('class_names')
        mean, std = blob.get('mean', mean), blob.get('std', std)
    elif dataset.startswith('synthetic') or dataset.endswith('-synthetic'):
        spec = dataset.replace('-synthetic', '').replace('synthetic', '').strip('-')
        if spec in SHAPES:
            C, T, H, n_train, n_test = SHAPES[spec]
        elif spec:
            C, n_train, T, H = (int(v) for v in spec.split('x'))
            n_test = max(1, n_train // 4)
        else:
            C, T, H, n_train, n_test = SHAPES['miniUCF101']
        xtr, ytr, xte, yte = _synthetic(C, n_train, n_test, T, H)
    else:
        raise NotImplementedError(
            f"get_dataset('{dataset}'): the video decoders of the reference are outside the B200 hot path; save the preloaded "
            f"tensors once as {path} (keys images_train/labels_train/images_test/labels_test) or use '<name>-synthetic'")
    num_classes = int(max(int(ytr.max()), int(yte.max())) + 1)
    if class_names is None:
        class_names = [str(c) for c in range(num_classes)]
    channel, im_size = int(xtr.shape[2]), (int(xtr.shape[3]), int(xtr.shape[4]))
    dst_train = _LabelledTensorDataset(xtr, ytr)
    dst_test = _LabelledTensorDataset(xte, yte)
    testloader = DataLoader(dst_test, batch_size=test_batch_size, shuffle=False, num_workers=num_workers)
    return channel, im_size, num_classes, class_names, mean, std, dst_train, dst_test, testloader
