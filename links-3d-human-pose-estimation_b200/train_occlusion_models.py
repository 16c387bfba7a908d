"""Drop-in for the reference's train_occlusion_models.py: same CLI flags and defaults (reference :27-44; only -t is
used by the step, -n only names the run), same step (OcclusionStep), optimiser and checkpoint names."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from links_b200.harness import add_common_args, train_occlusion  # noqa: E402

parser = argparse.ArgumentParser(description='Train 2D INN with PCA')
parser.add_argument("-n", "--num_bases", help="number of PCA bases", type=int, default=26)
parser.add_argument("-b", "--bl", help="bone lengths", type=float, default=50.0)
parser.add_argument("-t", "--translation", help="camera translation", type=float, default=10.0)
parser.add_argument("-r", "--rep2d", help="2d reprojection", type=float, default=1.0)
parser.add_argument("-o", "--rot3d", help="3d reconstruction", type=float, default=1.0)
parser.add_argument("-v", "--velocity", help="velocity", type=float, default=1.0)
parser.add_argument("-l", "--likelihood", help="likelihood", type=float, default=1.0)
add_common_args(parser, batch=256, epochs=10)

if __name__ == "__main__":
    train_occlusion(parser.parse_args())
