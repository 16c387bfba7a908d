"""Drop-in for the reference's eval_h36m.py (no flags there): lifts the test set with the left/right lifters, combines
with choice='right' and prints PA-MPJPE ('best' Procrustes) and N-MPJPE -- on the device, sharded across ranks."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from links_b200.harness import evaluate  # noqa: E402

parser = argparse.ArgumentParser(description="evaluate the left/right lifters")
g = parser.add_argument_group("links_b200 additions")
g.add_argument("--synthetic", type=int, default=1_000_000, help="number of synthetic test poses (no dataset is shipped)")
g.add_argument("--datafile", default=None, help="dataset pickle in the reference's format; subjects S9 / S11 are evaluated")
g.add_argument("--chunk", type=int, default=65536)
g.add_argument("--seed", type=int, default=0)
g.add_argument("--weights-dir", default="models")
g.add_argument("--random-init", action="store_true", help="seeded random lifters when the checkpoints are missing")

if __name__ == "__main__":
    evaluate(parser.parse_args())
