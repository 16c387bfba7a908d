"""Seeded synthetic H36M-shaped poses (17 joints) for benches and parity tests.

There is no dataset in the reference tree (``utils/h36m_dataset_class.py`` loads a pickle
that is not shipped), so the bench and tests use kinematic-tree skeletons shaped like what
``normalize_head`` (reference utils/helpers.py:198-207) produces: 2D rows are
``[B,34] = (17 x, 17 y)``, root-centred, mean root->head distance 0.1; 3D ground truth is
``[B,51] = (17 X, 17 Y, 17 Z)`` in millimetres in the camera frame.
"""
import numpy as np

# 16-bone tree of reference utils/helpers.py:140-141 (parent, child)
BONES = [(0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9), (9, 10), (8, 11), (11, 12),
         (12, 13), (8, 14), (14, 15), (15, 16)]
# relative bone lengths, reference train_left_right_lifter.py:76-79
BONE_REL = np.array([0.5180581, 1.73711136, 1.72285805, 0.5180552, 1.73710543, 1.72285651, 0.92087518,
                     0.98792375, 0.44812302, 0.44502545, 0.57462, 1.08121276, 0.9651687, 0.57461556,
                     1.08122523, 0.9651657])
# preferred direction per bone (image coords: +y down); legs down, spine up, arms sideways/down
_BIAS = np.array([[-1, 0, 0], [0, 1, 0], [0, 1, 0], [1, 0, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0],
                  [0, -1, 0], [0, -1, 0], [1, 0, 0], [0.3, 1, 0], [0, 1, 0], [-1, 0, 0], [-0.3, 1, 0], [0, 1, 0]],
                 dtype=np.float64)


def synth_poses(n, seed=1234, mean_bone_mm=250.0, cam_dist_mm=6000.0, dtype=np.float32):
    """Returns (poses_2d [n,34], poses_3d_mm [n,51])."""
    rng = np.random.RandomState(seed)
    lengths = BONE_REL[None, :] * rng.uniform(0.9, 1.1, size=(n, 16)) * mean_bone_mm
    dirs = rng.normal(size=(n, 16, 3)) * 0.6 + _BIAS[None] * 1.5
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    P = np.zeros((n, 17, 3))
    for b, (pa, ch) in enumerate(BONES):
        P[:, ch] = P[:, pa] + dirs[:, b] * lengths[:, b:b + 1]
    # random yaw so depth varies across joints
    yaw = rng.uniform(-np.pi, np.pi, size=n)
    c, s = np.cos(yaw), np.sin(yaw)
    X = c[:, None] * P[:, :, 0] + s[:, None] * P[:, :, 2]
    Z = -s[:, None] * P[:, :, 0] + c[:, None] * P[:, :, 2]
    P = np.stack((X, P[:, :, 1], Z), axis=2)
    z0 = cam_dist_mm * rng.uniform(0.9, 1.1, size=(n, 1))
    cam = P.copy()
    cam[:, :, 2] += z0
    p2d = cam[:, :, :2] / cam[:, :, 2:3]
    p2d = p2d - p2d[:, :1]
    scale = np.linalg.norm(p2d[:, 0] - p2d[:, 10], axis=1).mean()
    p2d = p2d / scale * 0.1
    poses_2d = np.concatenate((p2d[:, :, 0], p2d[:, :, 1]), axis=1)
    poses_3d = np.concatenate((cam[:, :, 0], cam[:, :, 1], cam[:, :, 2]), axis=1)
    return poses_2d.astype(dtype), poses_3d.astype(dtype)


def synth_pred_3d(poses_3d, seed=99, noise_mm=30.0, scale=0.01, mirror_frac=0.0):
    """A noisy, rescaled (and optionally mirrored) copy of GT poses: a stand-in prediction for metric tests."""
    rng = np.random.RandomState(seed)
    n = poses_3d.shape[0]
    p = poses_3d.reshape(n, 3, 17).astype(np.float64)
    p = p + rng.normal(size=p.shape) * noise_mm
    if mirror_frac > 0:
        m = rng.uniform(size=n) < mirror_frac
        p[m, 0] *= -1.0
    p = p * scale
    return p.reshape(n, 51).astype(poses_3d.dtype)
