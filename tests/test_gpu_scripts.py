"""GPU: the drop-in scripts (reference CLI preserved) run end to end on synthetic data: a few optimiser steps, the
reference's checkpoint file names / key sets, and evaluation from those checkpoints."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "links-3d-human-pose-estimation_b200")


def _run(script, *args):
    env = dict(os.environ, WANDB_MODE="disabled")
    r = subprocess.run([sys.executable, os.path.join(PKG, script)] + list(args), capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_pipeline_chains_through_reference_checkpoint_names(tmp_path):
    """The shipped pipeline in the reference's order, every stage reading what the previous one wrote (no --random-init):
    full-pose flow -> part flows -> both lifter trainers -> eval -> occlusion models (+ validation)."""
    wd = str(tmp_path)
    out = _run("train_full_pose_norm_flow.py", "-n", "34", "--synthetic", "1024", "--batch", "128", "--steps", "4", "--log-every", "2",
               "--weights-dir", wd)
    assert "step 4" in out
    sd = torch.load(os.path.join(wd, "norm_flow_sampling.pt"))
    assert len(sd) == 64 and sd["module_list.7.subnet.2.weight"].shape == (34, 1024)
    out = _run("train_leg_torso_left_right_norm_flow.py", "-l", "22", "--synthetic", "512", "--batch", "64", "--steps", "3",
               "--log-every", "1", "--weights-dir", wd)
    assert "step 3" in out and "dist_2d_torso=" in out
    sd = torch.load(os.path.join(wd, "mpi_norm_flow_legs_2.pt"))
    assert len(sd) == 64 and sd["module_list.0.subnet.2.weight"].shape == (14, 1024)
    out = _run("train_left_right_lifter.py", "-b", "50", "-t", "10", "--synthetic", "1024", "--batch", "128", "--steps", "6",
               "--log-every", "3", "--weights-dir", wd)
    assert "step 6" in out and "loss=" in out
    sd = torch.load(os.path.join(wd, "left_side_lifter_final.pt"))
    assert len(sd) == 62 and sd["downscale.weight"].shape == (11, 1024) and "res_pose2.bn1.weight" in sd
    out = _run("train_leg_torso_lifter.py", "--synthetic", "1024", "--batch", "128", "--steps", "4", "--log-every", "2",
               "--weights-dir", wd, "-l", "0.5", "--val", "777")
    assert "step 4" in out and "validation" in out and "pck=" in out
    assert os.path.exists(os.path.join(wd, "leg_lifter.pt")) and os.path.exists(os.path.join(wd, "torso_lifter.pt"))
    out = _run("eval_h36m.py", "--synthetic", "5000", "--chunk", "2048", "--weights-dir", wd)
    assert "PA-MPJPE:" in out and "N-MPJPE:" in out
    out = _run("train_occlusion_models.py", "-n", "26", "--synthetic", "512", "--batch", "64", "--steps", "3", "--log-every", "1",
               "--weights-dir", wd, "--val", "300")
    assert "step 3" in out and "validation" in out and "pa_torso=" in out
    sd = torch.load(os.path.join(wd, "occlusion_model_weights", "torso_estimator.pt"))
    assert len(sd) == 36 and sd["downscale.weight"].shape == (30, 1024)


def test_missing_checkpoint_raises_unless_random_init(tmp_path):
    """The reference's torch.load raises on a missing file; so do the drop-in scripts, unless --random-init."""
    env = dict(os.environ, WANDB_MODE="disabled")
    args = [sys.executable, os.path.join(PKG, "train_leg_torso_lifter.py"), "--synthetic", "256", "--batch", "64", "--steps", "1",
            "--weights-dir", str(tmp_path), "--no-save"]
    r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0 and "FileNotFoundError" in r.stderr
    r = subprocess.run(args + ["--random-init", "--log-every", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "step 1" in r.stdout, r.stderr[-2000:]


def test_scripts_on_a_dataset_pickle(tmp_path):
    """--datafile: the reference's pickle layout -> dataset drop-in -> sharded loader -> device prefetcher -> step."""
    import pickle
    import numpy as np
    sys.path.insert(0, PKG)
    from links_b200.synth import synth_poses
    data = {}
    for i, s in enumerate(['S1', 'S5', 'S6', 'S7', 'S8', 'S9', 'S11']):
        x2d, gt = synth_poses(300, seed=50 + i)
        data[s] = {"poses_2d": np.ascontiguousarray(x2d.reshape(-1, 2, 17).transpose(0, 2, 1)) * 1000.0,
                   "poses_3d": np.ascontiguousarray(gt.reshape(-1, 3, 17).transpose(0, 2, 1)).astype(np.float64)}
    path = str(tmp_path / "h36m_like.pkl")
    with open(path, "wb") as f:
        pickle.dump(data, f)
    wd = str(tmp_path)
    out = _run("train_left_right_lifter.py", "--datafile", path, "--batch", "128", "--steps", "5", "--log-every", "5",
               "--weights-dir", wd, "--random-init")
    assert "step 5" in out and "loss=" in out
    out = _run("eval_h36m.py", "--datafile", path, "--chunk", "1024", "--weights-dir", wd)     # reads the lifters saved above
    assert "PA-MPJPE:" in out and "N-MPJPE:" in out
