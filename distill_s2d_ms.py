"""Drop-in for the reference's distill_s2d_ms.py (same flags, :451-506): DM / MTT with static + dynamic memory on the
B200 kernels.  `python distill_s2d_ms.py --method DM --dataset miniUCF101-synthetic --vpc 1 --spc 2 --dpc 2 ...`;
`torchrun --nproc-per-node N distill_s2d_ms.py ...` shards the classes across N GPUs."""
from video_distillation_b200.cli import main_s2d as main, s2d_parser

if __name__ == '__main__':
    from video_distillation_b200.cli import init_distributed
    init_distributed()
    main(s2d_parser().parse_args())
