"""CPU oracle for the distillation inner loop (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference
algorithm for the hot path named in BASELINE.json:north_star.  It is the
*checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``video_distillation_b200`` never imports ``oracle``.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the live reference modules imported
from ``/root/reference`` by ``oracle/make_golden.py``; the resulting vectors are
committed under ``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.
"""
from .convnet3d import (  # noqa: F401
    init_convnet3d, convnet3d_features, convnet3d_embed, convnet3d_forward,
    convnet3d_param_names, flatten_params, unflatten_params, embed_dim, routing_codes,
)
from .composer import init_hallucinator, compose  # noqa: F401
from .dm import (  # noqa: F401
    sample_real_indices, s2d_sample_indices, dm_loss, dm_baseline_iteration,
    dm_s2d_iteration, sgd_momentum_step,
)
from .mtt import mtt_sample_step_indices, mtt_s2d_iteration, mtt_baseline_iteration  # noqa: F401
