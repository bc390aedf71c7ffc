"""Host-side plumbing of the precision modes (no CUDA): driver-level precision strings -> conv backend of the trio."""
import argparse

import pytest


def test_backend_for_precision_and_backend_switch():
    from video_distillation_b200 import ops
    assert ops.backend_for_precision('fp32') == 'fp32'
    assert ops.backend_for_precision('bf16') == 'tc'
    assert ops.backend_for_precision('bf16x3') == 'tc_x3'
    prev = ops.set_conv_backend('tc_x3')
    try:
        assert ops.set_conv_backend('tc') == 'tc_x3'
        with pytest.raises(ValueError):
            ops.set_conv_backend('cudnn')
    finally:
        ops.set_conv_backend(prev)
    assert ops.set_conv_backend(prev) == prev


@pytest.mark.parametrize('cli_precision,mtt', [('f16x3r2', 'bf16x3'), ('f16x3', 'bf16x3'), ('bf16x3', 'bf16x3'), ('bf16', 'bf16'), ('fp32', 'fp32')])
def test_mtt_precision_of_the_drivers(cli_precision, mtt):
    """The MTT unroll runs the parity-grade all-split trio unless the throughput mode or the exact kernels are asked for."""
    from video_distillation_b200 import cli
    assert cli._mtt_precision(argparse.Namespace(precision=cli_precision)) == mtt


def test_mtt_trainers_validate_precision():
    from video_distillation_b200.distill import MTTBaselineTrainer, MTTS2DTrainer
    for cls in (MTTS2DTrainer, MTTBaselineTrainer):
        with pytest.raises(ValueError):
            cls(num_classes=2, precision='tf32', device='cpu')


def test_xcol_cache_is_opt_in(monkeypatch):
    from video_distillation_b200 import tc_trio
    monkeypatch.delenv('VD_XCOL_CACHE', raising=False)
    tc_trio.xcol_cache_begin()
    assert tc_trio._xcol_cache is None
    tc_trio.xcol_cache_end()
    monkeypatch.setenv('VD_XCOL_CACHE', '1')
    tc_trio.xcol_cache_begin()
    assert tc_trio._xcol_cache == {}
    tc_trio.xcol_cache_end()
    assert tc_trio._xcol_cache is None
