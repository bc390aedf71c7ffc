"""Drop-in module name of the reference (`from utils import get_network, evaluate_synset, Conv3DNet, ...`,
distill_s2d_ms.py:10)."""
from video_distillation_b200.utils import (  # noqa: F401
    get_default_convnet_setting, get_network, get_time, get_eval_pool, Conv3DNet, TensorDataset,
    MultiStaticSharedDataset, epoch, evaluate_synset, get_dataset, get_loops, ParamDiffAug, DiffAugment, match_loss, get_daparam)
