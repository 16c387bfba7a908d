"""Integer index maps of the LInKs hot path (bit-exact gathers) and the geometry-kernel map struct.

Reference: utils/helpers.py:40-65,81-91,139-141; train_leg_torso_lifter.py:147-148;
train_occlusion_models.py:176-191.
"""
from . import _cabi

J = 17
RIGHT_JOINTS = [0, 1, 2, 3, 7, 8, 9, 10, 14, 15, 16]      # helpers.py:57-59
LEFT_JOINTS = [0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]       # helpers.py:61-63
LEG_JOINTS = list(range(0, 7))
TORSO_JOINTS = list(range(7, 17))
# helpers.py:40-53 -- per full-pose joint (side, index); side 0 = left part, 1 = right part
COMBINE_RIGHT = [(1, 0), (1, 1), (1, 2), (1, 3), (0, 1), (0, 2), (0, 3), (1, 4), (1, 5), (1, 6), (1, 7),
                 (0, 8), (0, 9), (0, 10), (1, 8), (1, 9), (1, 10)]
COMBINE_LEFT = [(0, 0), (1, 1), (1, 2), (1, 3), (0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (0, 7),
                (0, 8), (0, 9), (0, 10), (1, 8), (1, 9), (1, 10)]
# helpers.py:140-141 (parent, child) of the 16 bones
BONES = [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [0, 7], [7, 8], [8, 9], [9, 10], [8, 11], [11, 12], [12, 13],
         [8, 14], [14, 15], [15, 16]]
BONE_REL_H36M = [0.5180581, 1.73711136, 1.72285805, 0.5180552, 1.73710543, 1.72285651, 0.92087518, 0.98792375,
                 0.44812302, 0.44502545, 0.57462, 1.08121276, 0.9651687, 0.57461556, 1.08122523, 0.9651657]
BONE_REL_MPI = [0.48069107, 1.84637771, 1.49564841, 0.48069107, 1.84301997, 1.4956484, 0.90757932, 0.99706493,
                0.34679742, 0.69380255, 0.57843534, 1.20698327, 0.92306225, 0.5741528, 1.20698326, 0.92306223]


def part_index(joints, dims=2, stride=J):
    """Flat indices gathering `joints` of a [dims, 17] row into [dims, len(joints)] (x's, y's(, z's))."""
    return [a * stride + j for a in range(dims) for j in joints]


def split_lr_3d_index(joints):
    """split_data_left_right_3d (helpers.py:81-91): the [B,3,17] buffer is *viewed* as [1.5B,2,17], so the
    gather mixes row pairs.  Returns 2*33 offsets relative to a row pair (period 2)."""
    out = []
    for sub in range(2):
        for c in range(33):
            o = 33 * sub + c
            out.append(34 * (o // 22) + 17 * ((o % 22) // 11) + joints[o % 11])
    return out


# occlusion targets / inputs (train_occlusion_models.py:176-191) as joint lists over [3,17] rows
def _jl(*ranges):
    out = []
    for a, b in ranges:
        out += list(range(a, b))
    return out


OCC_NAMES = ("left_arm", "right_arm", "left_leg", "right_leg", "left_side", "right_side", "both_legs", "torso")
OCC_TARGET_JOINTS = {
    "left_arm": _jl((11, 14)), "right_arm": _jl((14, 17)), "left_leg": _jl((4, 7)), "right_leg": _jl((1, 4)),
    "left_side": _jl((4, 7), (11, 14)), "right_side": _jl((1, 4), (14, 17)), "both_legs": _jl((1, 7)),
    "torso": _jl((7, 17)),
}
OCC_INPUT_JOINTS = {
    "left_arm": _jl((0, 11), (14, 17)), "right_arm": _jl((0, 14)), "left_leg": _jl((0, 4), (7, 17)),
    "right_leg": _jl((0, 1), (4, 17)), "torso": _jl((0, 7)), "both_legs": _jl((0, 1), (7, 17)),
}


def occ_target_index(name):
    return part_index(OCC_TARGET_JOINTS[name], dims=3)


def occ_input_index(name):
    """(index list, period).  left_side / right_side use the scrambled pair gather (:191):
    (no_right_side, no_left_side) = split_data_left_right_3d -> (left-list gather, right-list gather);
    left_predictor takes no_left_side (= right-list gather)."""
    if name == "left_side":
        return split_lr_3d_index(RIGHT_JOINTS), 2
    if name == "right_side":
        return split_lr_3d_index(LEFT_JOINTS), 2
    return part_index(OCC_INPUT_JOINTS[name], dims=3), 1


def geom_maps(kind, cfg=None):
    """LinksGeomMaps for kind in {'lt', 'lr'}; cfg = dict(depth, weight_*)."""
    base = dict(depth=10.0, weight_bl=50.0, weight_2d=1.0, weight_3d=1.0, weight_likeli=1.0, weight_velocity=1.0)
    base.update(cfg or {})
    cfg = base
    m = _cabi.GeomMaps()
    if kind == "lt":
        m.V = 1
        m.n_joints[0], m.n_joints[1] = 7, 10
        for j in range(J):
            net = 0 if j < 7 else 1
            idx = j if j < 7 else j - 7
            m.src_net[0][j] = net
            m.src_net[1][j] = net
            m.col[j] = idx
            m.part_net[0][j] = net
            m.part_idx[0][j] = idx
            m.part_net[1][j] = -1
            m.part_idx[1][j] = 0
        bones = cfg.get("bone_rel", BONE_REL_MPI)      # train_leg_torso_lifter.py:97-100
    elif kind == "lr":
        m.V = 2
        m.n_joints[0], m.n_joints[1] = 11, 11
        for j in range(J):
            (sl, il), (sr, ir) = COMBINE_LEFT[j], COMBINE_RIGHT[j]
            assert il == ir
            m.src_net[0][j] = sl            # variant 0 = 'left' choice
            m.src_net[1][j] = sr            # variant 1 = 'right' choice
            m.col[j] = il
            # left lifter/flow read the left joints of the left-choice pose; right ones the right joints of
            # the right-choice pose (train_left_right_lifter.py:329-330,355-356)
            m.part_net[0][j] = 0 if j in LEFT_JOINTS else -1
            m.part_idx[0][j] = LEFT_JOINTS.index(j) if j in LEFT_JOINTS else 0
            m.part_net[1][j] = 1 if j in RIGHT_JOINTS else -1
            m.part_idx[1][j] = RIGHT_JOINTS.index(j) if j in RIGHT_JOINTS else 0
        bones = cfg.get("bone_rel", BONE_REL_H36M)     # train_left_right_lifter.py:76-79
    else:
        raise ValueError("kind must be 'lt' or 'lr'")
    for b in range(16):
        m.bone_rel[b] = bones[b]
    m.depth = cfg["depth"]
    m.w_likeli, m.w_2d, m.w_3d = cfg["weight_likeli"], cfg["weight_2d"], cfg["weight_3d"]
    m.w_vel, m.w_bl = cfg["weight_velocity"], cfg["weight_bl"]
    return m


def part_joint_lists(kind):
    return (LEG_JOINTS, TORSO_JOINTS) if kind == "lt" else (LEFT_JOINTS, RIGHT_JOINTS)
