"""Drop-in module name of the reference (`from networks import ConvNet3D`, networks.py:727)."""
from video_distillation_b200.networks import *  # noqa: F401,F403
from video_distillation_b200.networks import ConvNet3D  # noqa: F401
