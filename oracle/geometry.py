"""Oracle (test infrastructure): index maps, rotations, projection, bone lengths.

Restates reference ``utils/helpers.py`` and ``utils/rotation_conversions.py``.
Index maps are integer gathers and must be bit-exact.
"""
import numpy as np
import torch

J = 17
# helpers.py:55-65
RIGHT_JOINTS = [0, 1, 2, 3, 7, 8, 9, 10, 14, 15, 16]
LEFT_JOINTS = [0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]
LEG_JOINTS = list(range(0, 7))       # train_leg_torso_lifter.py:147
TORSO_JOINTS = list(range(7, 17))    # train_leg_torso_lifter.py:148
# helpers.py:140-141
BONES = [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [0, 7], [7, 8], [8, 9], [9, 10], [8, 11], [11, 12],
         [12, 13], [8, 14], [14, 15], [15, 16]]
# helpers.py:40-53: per full-pose joint, (side, index) with side 0 = left part, 1 = right part
COMBINE_RIGHT = [(1, 0), (1, 1), (1, 2), (1, 3), (0, 1), (0, 2), (0, 3), (1, 4), (1, 5), (1, 6), (1, 7),
                 (0, 8), (0, 9), (0, 10), (1, 8), (1, 9), (1, 10)]
COMBINE_LEFT = [(0, 0), (1, 1), (1, 2), (1, 3), (0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (0, 7),
                (0, 8), (0, 9), (0, 10), (1, 8), (1, 9), (1, 10)]
# train_left_right_lifter.py:76-79 (H36M) and train_leg_torso_lifter.py:97-100 (MPI)
BONE_REL_H36M = [0.5180581, 1.73711136, 1.72285805, 0.5180552, 1.73710543, 1.72285651, 0.92087518, 0.98792375,
                 0.44812302, 0.44502545, 0.57462, 1.08121276, 0.9651687, 0.57461556, 1.08122523, 0.9651657]
BONE_REL_MPI = [0.48069107, 1.84637771, 1.49564841, 0.48069107, 1.84301997, 1.4956484, 0.90757932, 0.99706493,
                0.34679742, 0.69380255, 0.57843534, 1.20698327, 0.92306225, 0.5741528, 1.20698326, 0.92306223]


def part_2d(data, joints):
    """Gather a joint subset of [M,34]=(17 x,17 y) into [M,2*len] = (x's, y's)."""
    d = data.reshape(-1, 2, J)
    return d[:, :, joints].reshape(-1, 2 * len(joints))


def split_data_left_right(data):
    """helpers.py:55-65 -> (left, right)."""
    return part_2d(data, LEFT_JOINTS), part_2d(data, RIGHT_JOINTS)


def split_data_left_right_3d(data):
    """helpers.py:81-91.  NOTE the reference reshapes a [B,3,17] tensor as [-1,2,17]:
    a scrambled gather mixing consecutive pose pairs (B must be even)."""
    d = data.reshape(-1, 2, J)
    right = d[:, :, RIGHT_JOINTS].reshape(-1, 33)
    left = d[:, :, LEFT_JOINTS].reshape(-1, 33)
    return left, right


def combine_left_right_1d(left, right, choice):
    """helpers.py:40-53 -> [M,17]."""
    table = COMBINE_RIGHT if choice == "right" else COMBINE_LEFT
    parts = (left.reshape(-1, 11), right.reshape(-1, 11))
    return torch.stack([parts[s][:, i] for s, i in table], dim=1)


def combine_left_right_nd(left, right, choice, dims):
    """helpers.py:7-38 (2d / 3d variants) -> [M, dims*17]."""
    table = COMBINE_RIGHT if choice == "right" else COMBINE_LEFT
    parts = (left.reshape(-1, dims, 11), right.reshape(-1, dims, 11))
    return torch.stack([parts[s][:, :, i] for s, i in table], dim=2).reshape(-1, dims * J)


def rot_x(a):
    """_axis_angle_rotation('X'), rotation_conversions.py:11-36.  a: [M] -> [M,3,3]."""
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones_like(a), torch.zeros_like(a)
    return torch.stack((o, z, z, z, c, -s, z, s, c), -1).reshape(a.shape + (3, 3))


def rot_y(a):
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones_like(a), torch.zeros_like(a)
    return torch.stack((c, z, s, z, o, z, -s, z, c), -1).reshape(a.shape + (3, 3))


def rot_z(a):
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones_like(a), torch.zeros_like(a)
    return torch.stack((c, -s, z, s, c, z, z, z, o), -1).reshape(a.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    """rotation_conversions.py:39-61."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError("Invalid convention %s." % convention)
    fn = {"X": rot_x, "Y": rot_y, "Z": rot_z}
    for letter in convention:
        if letter not in fn:
            raise ValueError("Invalid letter %s in convention string." % letter)
    m = [fn[c](euler_angles[..., i]) for i, c in enumerate(convention)]
    return m[0] @ m[1] @ m[2]


def perspective_projection(pose_3d):
    """helpers.py:262-267."""
    pose_3d = pose_3d.reshape(-1, 51)
    p2d = pose_3d[:, 0:34].reshape(-1, 2, J) / pose_3d[:, 34:51].reshape(-1, 1, J)
    return p2d.reshape(-1, 34)


def get_bone_lengths_all(poses):
    """helpers.py:139-151 -> [M,16]."""
    p = poses.reshape(-1, 3, J)
    a = p[:, :, [b[0] for b in BONES]]
    b = p[:, :, [b[1] for b in BONES]]
    return torch.norm(a - b, p=2, dim=1)


def lift(u, depth):
    """[x*d, y*d, d] as [M,3,17]  (train_leg_torso_lifter.py:188-190)."""
    uu = u.reshape(-1, 2, J)
    d = depth.reshape(-1, 1, J)
    return torch.cat((uu * d, d), dim=1)


def add_noise(latent, noise, noise_factor):
    """helpers.py:298-308 with the N(0,1) draw passed in."""
    return latent + noise_factor * (noise * latent)


def normalize_head(poses_2d, root_joint=0):
    """helpers.py:198-207 (numpy; mutates its argument exactly like the reference)."""
    p2d = poses_2d.reshape(-1, 2, J)
    p2d -= p2d[:, :, [root_joint]]
    scale = np.linalg.norm(p2d[:, :, 0] - p2d[:, :, 10], axis=1, keepdims=True)
    return poses_2d / scale.mean() * (1 / 10)


def normalize_head_test(poses_2d, scale=145.40964):
    """helpers.py:222-230."""
    p2d = poses_2d.reshape(-1, 2, J)
    p2d -= p2d[:, :, [0]]
    return poses_2d / scale * (1 / 10)
