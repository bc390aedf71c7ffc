"""Drop-in for the reference's distill_baseline.py (same flags, :366-417): DM / MTT on leaf synthetic videos on the
B200 kernels (the DC method is outside the hot path)."""
from video_distillation_b200.cli import main_baseline as main, baseline_parser

if __name__ == '__main__':
    from video_distillation_b200.cli import init_distributed
    init_distributed()
    main(baseline_parser().parse_args())
