"""Drop-in for the reference's train_leg_torso_lifter.py: same CLI flags and defaults (reference :23-37), same step
(LifterStep 'lt'), optimiser (Adam 2e-4, wd 1e-5, ExponentialLR 0.95 per epoch) and checkpoint names; runs on B200
through links_b200.  Launch with torchrun for data-parallel training.  wandb logging is not part of the hot path."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from links_b200.harness import add_common_args, train_lifters  # noqa: E402

parser = argparse.ArgumentParser(description='Train 2D INN with PCA')
parser.add_argument("-b", "--bl", help="bone lengths", type=float, default=50.0)
parser.add_argument("-t", "--translation", help="camera translation", type=float, default=10.0)
parser.add_argument("-r", "--rep2d", help="2d reprojection", type=float, default=1.0)
parser.add_argument("-o", "--rot3d", help="3d reconstruction", type=float, default=1.0)
parser.add_argument("-v", "--velocity", help="velocity", type=float, default=1.0)
parser.add_argument("-l", "--likelihood", help="likelihood", type=float, default=1.0)
add_common_args(parser, batch=256, epochs=100)

if __name__ == "__main__":
    train_lifters("lt", parser.parse_args())
