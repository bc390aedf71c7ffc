"""Drop-in for the reference's buffer.py (same flags, :107-131): trains expert ConvNet3D trajectories on the B200 kernels and
writes replay_buffer_{n}.pt in the reference's on-disk format."""
from video_distillation_b200.cli import main_buffer as main, buffer_parser

if __name__ == '__main__':
    main(buffer_parser().parse_args())
