"""Drop-in module name of the reference (`from reparam_module import ReparamModule`, distill_s2d_ms.py:14)."""
from video_distillation_b200.reparam_module import ReparamModule  # noqa: F401
