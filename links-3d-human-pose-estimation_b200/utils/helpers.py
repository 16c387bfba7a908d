"""Drop-in for reference ``utils/helpers.py``: the same function names and tensor shapes.

The joint-subset maps are integer gathers driven by the index tables of ``links_b200.maps`` (bit-exact; the
fused step kernels use the very same tables).  These host-side helpers are device-agnostic tensor code for
callers outside the fused training/eval steps (data preparation, scripts, notebooks).
"""
import random

import numpy as np
import torch
import torch.nn as nn

from links_b200 import maps as _M

_J = 17


def _idx(joints, device):
    return torch.as_tensor(joints, dtype=torch.long, device=device)


def _gather_joints(data, dims, joints):
    d = data.reshape(-1, dims, _J)
    return d.index_select(2, _idx(joints, d.device))


def _combine(left_split, right_split, choice, dims):
    table = _M.COMBINE_RIGHT if choice == 'right' else _M.COMBINE_LEFT
    both = torch.cat((left_split.reshape(-1, dims, 11), right_split.reshape(-1, dims, 11)), dim=2)   # [M,dims,22]
    return both.index_select(2, _idx([side * 11 + i for side, i in table], both.device))


def combine_left_right_pred_3d(left_split, right_split, choice):
    return _combine(left_split, right_split, choice, 3).reshape(-1, 51)


def combine_left_right_pred_2d(left_split, right_split, choice):
    return _combine(left_split, right_split, choice, 2).reshape(-1, 34)


def combine_left_right_pred_1d(left_split, right_split, choice):
    return _combine(left_split, right_split, choice, 1)


def split_data_left_right(data):
    """-> (left, right), each [M, 22] = (11 x, 11 y)."""
    return (_gather_joints(data, 2, _M.LEFT_JOINTS).reshape(-1, 22), _gather_joints(data, 2, _M.RIGHT_JOINTS).reshape(-1, 22))


def split_data_left_right_v2(data):
    right = [0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 13]
    left = [0, 4, 5, 6, 7, 8, 9, 10, 14, 15, 16]
    return _gather_joints(data, 2, left).reshape(-1, 22), _gather_joints(data, 2, right).reshape(-1, 22)


def split_data_left_right_3d(data):
    """The reference views the [B,3,17] buffer as [-1,2,17] before gathering (a pair-mixing scramble that
    defines parity; B must be even)."""
    return (_gather_joints(data, 2, _M.LEFT_JOINTS).reshape(-1, 33), _gather_joints(data, 2, _M.RIGHT_JOINTS).reshape(-1, 33))


def split_data_left_right_numpy(data):
    d = data.reshape(-1, 2, _J)
    return d[:, :, _M.LEFT_JOINTS].reshape(-1, 22), d[:, :, _M.RIGHT_JOINTS].reshape(-1, 22)


def temporal_split_data_left_right(data):
    d = data.reshape(-1, 2, 2, _J)
    left = d.index_select(3, _idx(_M.LEFT_JOINTS, d.device)).reshape(-1, 44)
    right = d.index_select(3, _idx(_M.RIGHT_JOINTS, d.device)).reshape(-1, 44)
    return left, right


def combine_left_right_occluded_3d(occluded_part, visible_part, part_occluded):
    occ = occluded_part.reshape(-1, 3, 6)
    vis = visible_part.reshape(-1, 3, 11)
    both = torch.cat((vis, occ), dim=2)          # 0..10 visible, 11..16 occluded
    if part_occluded == 'right':
        order = [0, 11, 12, 13, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 15, 16]
    else:
        order = [0, 1, 2, 3, 11, 12, 13, 4, 5, 6, 7, 14, 15, 16, 8, 9, 10]
    return both.index_select(2, _idx(order, both.device))


def _bone_lengths(poses, n_joints, bones):
    p = poses.reshape(-1, 3, n_joints)
    a = p.index_select(2, _idx([b[0] for b in bones], p.device))
    b = p.index_select(2, _idx([b[1] for b in bones], p.device))
    return torch.norm(a - b, p=2, dim=1)


def get_bone_lengths_all(poses):
    return _bone_lengths(poses, 17, _M.BONES)


def get_bone_lengths_legs(poses):
    return _bone_lengths(poses, 7, _M.BONES[:6])


def get_bone_lengths_torso(poses):
    p = poses.reshape(-1, 3, 10)
    p = torch.cat((torch.zeros(p.shape[0], 3, 1, dtype=p.dtype, device=p.device), p), dim=2)
    return _bone_lengths(p, 11, [[0, 1], [1, 2], [2, 3], [3, 4], [2, 5], [5, 6], [6, 7], [2, 8], [8, 9], [9, 10]])


def get_bone_lengths_left_right(poses):
    return _bone_lengths(poses, 11, [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [6, 7], [5, 8], [8, 9], [9, 10]])


def _root_centre_inplace(poses_2d, root_joint=0):
    p2d = poses_2d.reshape(-1, 2, _J)            # a view: the reference mutates its argument the same way
    p2d -= p2d[:, :, [root_joint]]
    return p2d


def normalize_head(poses_2d, root_joint=0):
    p2d = _root_centre_inplace(poses_2d, root_joint)
    scale = np.linalg.norm(p2d[:, :, 0] - p2d[:, :, 10], axis=1, keepdims=True)
    return poses_2d / scale.mean() * (1 / 10)


def _normalize_fixed(poses_2d, scale):
    _root_centre_inplace(poses_2d, 0)
    return poses_2d / scale * (1 / 10)


def normalize_head_test(poses_2d, scale=145.40964):
    return _normalize_fixed(poses_2d, scale)


def normalize_head_test_mpi_chest(poses_2d, scale=318.79249520730474):
    return _normalize_fixed(poses_2d, scale)


def normalize_head_test_mpi_vnect(poses_2d, scale=302.8530630720979):
    return _normalize_fixed(poses_2d, scale)


def normalize_head_test_temporal(poses_2d, scale=145.40419):
    return _normalize_fixed(poses_2d, scale)


def interpolate_gaussian_batch(latent_variables, t):
    if len(latent_variables) % 2 != 0:
        raise ValueError("Batch size must be even for interpolation.")
    pairs = latent_variables.reshape(-1, 2, 34)
    return (1 - t) * pairs[:, 0] + t * pairs[:, 1]


def _project(pose_3d, nj):
    p = pose_3d.reshape(-1, 3 * nj)
    return (p[:, :2 * nj].reshape(-1, 2, nj) / p[:, 2 * nj:].reshape(-1, 1, nj)).reshape(-1, 2 * nj)


def perspective_projection(pose_3d):
    return _project(pose_3d, 17)


def perspective_projection_legs(pose_3d):
    return _project(pose_3d, 7)


def perspective_projection_torso(pose_3d):
    return _project(pose_3d, 10)


def perspective_projection_left_right(pose_3d):
    return _project(pose_3d, 11)


def subnet_fc(dims_in, dims_out):
    return nn.Sequential(nn.Linear(dims_in, 1024), nn.ReLU(), nn.Linear(1024, dims_out))


def add_noise(latent_vars, noise_factor):
    noise = torch.randn_like(latent_vars)
    return latent_vars + noise_factor * (noise * latent_vars)


def occlusion_create(poses_2d):
    """Zero a random suffix of the left-leg keypoints of every pose (only 'left_leg' is enabled in the reference)."""
    out = poses_2d.clone().reshape(-1, 2, _J)
    limbs = {'left_leg': [[6], [5, 6], [4, 5, 6]], 'right_leg': [[3], [2, 3], [1, 2, 3]],
             'left_arm': [[11], [11, 12], [11, 12, 13]], 'right_arm': [[14], [14, 15], [14, 15, 16]]}
    for i in range(len(out)):
        limb = random.choice([['left_leg']])
        key = [k for k in ('left_leg', 'right_leg', 'left_arm') if k in limb]
        kps = random.choice(limbs[key[0]] if key else limbs['right_arm'])
        for kp in kps:
            out[i, :, kp] = 0.0
    return out.reshape(-1, 34)
