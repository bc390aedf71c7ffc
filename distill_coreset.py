"""Drop-in for the reference's distill_coreset.py (same flags, :148-167): k-center / herding coresets over ConvNet3D embeddings
computed on the B200 kernels."""
from video_distillation_b200.cli import main_coreset as main, coreset_parser

if __name__ == '__main__':
    main(coreset_parser().parse_args())
