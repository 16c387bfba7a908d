"""Occlusion inference assembly (reference train_occlusion_models.py:316-389, its validation_step): from the four part
lifters' depths build the root-centred 3D parts, form the eight "visible part" predictor inputs, and scatter each
predictor's output back into a full 17-joint pose.  Everything here is integer gather / concatenation on [M, 3, k]
tensors (device-agnostic torch indexing), described by ONE table instead of eight hand-written torch.cat chains."""
import torch

from . import maps

PART_JOINTS = {"legs": maps.LEG_JOINTS, "torso": maps.TORSO_JOINTS, "left": maps.LEFT_JOINTS, "right": maps.RIGHT_JOINTS}

# case -> (predictor name, visible pieces [(part, first, last)], predicted joints)
#   reference :364-371 (inputs), :382-389 (assembly), combine_pose_and_limb :67-78, utils/helpers.py:121-136
CASES = {
    "la": ("left_arm", [("legs", 0, 7), ("right", 4, 11)], [11, 12, 13]),
    "ra": ("right_arm", [("legs", 0, 7), ("left", 4, 11)], [14, 15, 16]),
    "ll": ("left_leg", [("right", 0, 4), ("torso", 0, 10)], [4, 5, 6]),
    "rl": ("right_leg", [("left", 0, 4), ("torso", 0, 10)], [1, 2, 3]),
    "torso": ("torso", [("legs", 0, 7)], list(range(7, 17))),
    "legs": ("both_legs", [("legs", 0, 1), ("torso", 0, 10)], list(range(1, 7))),
    "left": ("left_side", [("right", 0, 11)], [4, 5, 6, 11, 12, 13]),
    "right": ("right_side", [("left", 0, 11)], [1, 2, 3, 14, 15, 16]),
}


def visible_joints(case):
    return [PART_JOINTS[p][i] for p, a, b in CASES[case][1] for i in range(a, b)]


def full_pose_permutation(case):
    """Position of full-pose joint j in cat(visible joints, predicted joints)."""
    order = visible_joints(case) + CASES[case][2]
    assert sorted(order) == list(range(17)), (case, order)
    return [order.index(j) for j in range(17)]


def lift_parts(poses_2d, legs_pred, torso_pred, left_pred, right_pred, depth):
    """:327-361 -> root-centred parts {legs [M,3,7], torso [M,3,10], left [M,3,11], right [M,3,11]}.  The root depth
    offsets are forced to 0 before adding `depth`; the torso is centred on the LEG part's root (:358)."""
    x = poses_2d.reshape(-1, 2, 17)
    d = {"legs": legs_pred.clone(), "torso": torso_pred.clone(), "left": left_pred.clone(), "right": right_pred.clone()}
    for k in ("legs", "left", "right"):
        d[k][:, 0] = 0.0
    parts = {}
    for k, joints in PART_JOINTS.items():
        dk = (d[k] + depth).unsqueeze(1)                       # [M,1,k]
        parts[k] = torch.cat((x[:, :, joints] * dk, dk), dim=1)
    root_legs = parts["legs"][:, :, :1].clone()
    parts["torso"] = parts["torso"] - root_legs
    for k in ("legs", "left", "right"):
        parts[k] = parts[k] - parts[k][:, :, :1]
    return parts


def visible_input(parts, case):
    """The predictor input of one occlusion case: [M, 3*k] (coordinate-major)."""
    pieces = [parts[p][:, :, a:b] for p, a, b in CASES[case][1]]
    v = torch.cat(pieces, dim=2) if len(pieces) > 1 else pieces[0]
    return v.reshape(v.shape[0], -1)


def assemble(case, visible, pred):
    """visible [M,3k], pred [M,3(17-k)] -> full pose [M,51]."""
    M = visible.shape[0]
    both = torch.cat((visible.reshape(M, 3, -1), pred.reshape(M, 3, -1)), dim=2)
    idx = torch.tensor(full_pose_permutation(case), dtype=torch.long, device=visible.device)
    return both.index_select(2, idx).reshape(M, 51)


def combine_pose_and_limb(pose, limb, which_limb):
    """Drop-in for the function of the same name in the reference script (:67-78)."""
    if which_limb not in ("ll", "rl", "la", "ra"):
        raise ValueError("which_limb must be one of 'll', 'rl', 'la', 'ra'")
    return assemble(which_limb, pose.reshape(-1, 42), limb.reshape(-1, 9))


def to_global(full_pose, depth):
    """:391-398: camera-frame pose, z += depth."""
    out = full_pose.clone()
    out[:, 34:51] += depth
    return out


class OcclusionValidator:
    """validation_step of the occlusion script (:316-509) without its host loops: lift the four parts, run the eight
    predictors on their visible-part inputs, assemble the eight full poses and score each against the ground truth
    (PA-MPJPE 'best' and scale-matched MPJPE, batched on the device instead of 8 numpy calls per pose).

    lifters:    {"legs" | "torso" | "left" | "right": callable(x[M,2k]) -> (depth offsets [M,k], angle)}
    predictors: {maps.OCC_NAMES: callable(x[M,3k]) -> [M,3(17-k)]}
    metrics:    object with pmpjpe_best(gt[M,51], pred[M,51]) and mpjpe(gt, pred, num_joints=17, root_joint=0),
                both returning per-pose errors (utils.metrics_batch.Metrics on the GPU)."""

    def __init__(self, lifters, predictors, metrics, depth=10.0):
        self.lifters, self.predictors, self.metrics, self.depth = lifters, predictors, metrics, depth

    @classmethod
    def from_params(cls, lifter_params, predictor_params, depth=10.0, device="cuda"):
        """Drop-in modules (utils.models_def) on the device, loaded from reference-style state dicts."""
        from utils import models_def as MD
        from utils.metrics_batch import Metrics
        lift_cls = {"legs": (MD.Leg_Lifter, 7), "torso": (MD.Torso_Lifter, 10), "left": (MD.Left_Right_Lifter, 11),
                    "right": (MD.Left_Right_Lifter, 11)}
        pred_cls = {"left_arm": MD.Occluded_Limb_Predictor, "right_arm": MD.Occluded_Limb_Predictor,
                    "left_leg": MD.Occluded_Limb_Predictor, "right_leg": MD.Occluded_Limb_Predictor,
                    "left_side": MD.Occluded_Left_Right_Predictor, "right_side": MD.Occluded_Left_Right_Predictor,
                    "both_legs": MD.Occluded_Legs_Predictor, "torso": MD.Occluded_Torso_Predictor}
        lifters, predictors = {}, {}
        for k, (c, nj) in lift_cls.items():
            m = c(use_batchnorm=False, num_joints=nj, use_dropout=False, d_rate=0.25).to(device)
            m.load_state_dict(lifter_params[k], strict=False)
            lifters[k] = m.eval()
        for case, (name, _, _) in CASES.items():
            m = pred_cls[name](use_batchnorm=False, num_joints=len(visible_joints(case))).to(device)
            m.load_state_dict(predictor_params[name], strict=False)
            predictors[name] = m.eval()
        return cls(lifters, predictors, Metrics(), depth)

    def load_predictors(self, predictor_params):
        """Refresh the predictors from a model that is still training (epoch-end validation)."""
        for name, m in self.predictors.items():
            m.load_state_dict(predictor_params[name], strict=False)

    @torch.no_grad()
    def run(self, poses_2d, gt_3d):
        """poses_2d [M,34], gt_3d [M,51] -> {"pa_<case>", "mpjpe_scaled_<case>"} means over the M poses (floats)."""
        x = poses_2d.reshape(-1, 2, 17)
        depths = {k: self.lifters[k](x[:, :, j].reshape(x.shape[0], -1).contiguous())[0] for k, j in PART_JOINTS.items()}
        parts = lift_parts(poses_2d, depths["legs"], depths["torso"], depths["left"], depths["right"], self.depth)
        out = {}
        for case, (name, _, _) in CASES.items():
            v = visible_input(parts, case).contiguous()
            full = to_global(assemble(case, v, self.predictors[name](v)), self.depth).contiguous()
            out["pa_" + case] = float(self.metrics.pmpjpe_best(gt_3d, full).double().mean())
            out["mpjpe_scaled_" + case] = float(self.metrics.mpjpe(gt_3d, full, num_joints=17, root_joint=0).double().mean())
        return out
