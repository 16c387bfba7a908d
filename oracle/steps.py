"""Oracle (test infrastructure): step bodies of the reference training/eval scripts.

The scripts themselves cannot be imported (module-level argparse / wandb.init /
torch.load of absent weights / pytorch_lightning classes), so the bodies are restated
here line range by line range on top of ``oracle.{nets,flow,geometry,metrics}``.
All randomness is passed in as tensors (CPU and CUDA generators can never agree).

Every step is a pure function ``(inputs, params, draws, cfg) -> dict of losses``; gradients
come from torch autograd on the returned ``loss``; parameter updates from
``torch.optim.Adam`` (the reference's optimiser, train_leg_torso_lifter.py:111-114).
"""
import numpy as np
import torch

from . import flow as F
from . import geometry as G
from . import metrics as M
from . import nets as N

DEFAULT_CFG = dict(depth=10.0, weight_bl=50.0, weight_2d=1.0, weight_3d=1.0, weight_likeli=1.0,
                   weight_velocity=1.0)  # argparse defaults, train_leg_torso_lifter.py:23-35


def sample_poses(x, full_flow, noise):
    """no_grad sampling block, train_leg_torso_lifter.py:133-142 -> [2B,34]."""
    with torch.no_grad():
        z, _ = F.inn_forward(x, full_flow)
        zn = G.add_noise(z, noise, 0.2)
        s, _ = F.inn_forward(zn, full_flow, rev=True)
        s = s.reshape(-1, 2, 17).clone()
        s[:, :, [0]] = 0.0
        s = s.reshape(-1, x.shape[1])
        return torch.cat((x, s), dim=0)


def _rotation(props, eps_x, u_y):
    """train_leg_torso_lifter.py:159-181 with use_elevation=True."""
    n = props.shape[0]
    zeros = torch.zeros((n, 1), dtype=props.dtype)
    R_comp = G.euler_angles_to_matrix(torch.cat((props, zeros, zeros), dim=1), "XYZ")
    elevation = torch.cat((props.mean().reshape(1), props.std().reshape(1)))
    x_ang = (-elevation[0]) + elevation[1] * eps_x.reshape(n, 1)
    y_ang = (u_y.reshape(n, 1) - 0.5) * 1.99 * np.pi
    Rx = G.euler_angles_to_matrix(torch.cat((x_ang, zeros, zeros), dim=1), "XYZ")
    Ry = G.euler_angles_to_matrix(torch.cat((zeros, y_ang, zeros), dim=1), "XYZ")
    return Rx @ (Ry @ R_comp)


def _lift_centered(u, pred, depth, clamp=True):
    """:185-192 -- depth offset, clamp (in place on a copy: zero grad where clamped), lift, root-centre."""
    d = pred + depth
    if clamp:
        d = torch.where(d < 1.0, torch.ones_like(d), d)
    p3 = G.lift(u, d)
    return p3 - p3[:, :, [0]]


def _rotate_project(R, p3, depth):
    """:195-199 -> (rot_poses [N,51], rot_2d [N,34])."""
    rot = (R @ p3).reshape(-1, 51)
    glob = torch.cat((rot[:, 0:34], rot[:, 34:51] + depth), dim=1)
    return rot, G.perspective_projection(glob)


def _consistency_terms(u, R, p3, rot, rot_2d, pred_rot, bone_rel, depth):
    """:231-259 for one (pose variant): returns L3d, rep_rot, re_rot_3d, bl_prior, re_rot (S)."""
    d2 = pred_rot + depth
    d2 = torch.where(d2 < 1.0, torch.ones_like(d2), d2)
    p3r = G.lift(rot_2d, d2)
    p3r = p3r - p3r[:, :, [0]]
    L3d = (rot - p3r.reshape(-1, 51)).norm(dim=1).mean()
    re_rot = (R.permute(0, 2, 1) @ p3r).reshape(-1, 51)
    glob = torch.cat((re_rot[:, 0:34], re_rot[:, 34:51] + depth), dim=1)
    re_2d = G.perspective_projection(glob)
    rep_rot = (re_2d - u).abs().sum(dim=1).mean()
    n_pairs = p3.shape[0] // 2
    pp = p3[0:2 * n_pairs].reshape(2 * n_pairs, 51).reshape(-1, 2, 51)
    pr = re_rot[0:2 * n_pairs].reshape(-1, 2, 51)
    re_rot_3d = ((pp[:, 0] - pp[:, 1]) - (pr[:, 0] - pr[:, 1])).norm(dim=1).mean()
    bl = G.get_bone_lengths_all(p3.reshape(-1, 51))
    rel = bl / bl.mean(dim=1, keepdim=True)
    bl_prior = (bone_rel - rel).square().sum(dim=1).mean()
    return L3d, rep_rot, re_rot_3d, bl_prior


def _total(losses, cfg):
    """:266-272."""
    return (cfg["weight_likeli"] * losses["likeli"] + cfg["weight_2d"] * losses["rep_rot"]
            + cfg["weight_3d"] * losses["L3d"] + cfg["weight_velocity"] * losses["re_rot_3d"]
            + cfg["weight_bl"] * losses["bl_prior"])


def lt_step(u, leg, torso, leg_flow, torso_flow, eps_x, u_y, cfg=DEFAULT_CFG, bone_rel=None, aux=None):
    """Leg/torso lifter training step, train_leg_torso_lifter.py:147-272 (after sampling).

    u: [N,34] (= cat(real, sampled)); eps_x ~ N(0,1) [N]; u_y ~ U(0,1) [N]."""
    depth = cfg["depth"]
    if bone_rel is None:
        bone_rel = torch.tensor(G.BONE_REL_MPI, dtype=u.dtype)  # :97-100
    legs_pred, legs_angle = N.lifter_forward(G.part_2d(u, G.LEG_JOINTS), leg)
    torso_pred, torso_angle = N.lifter_forward(G.part_2d(u, G.TORSO_JOINTS), torso)
    props = (legs_angle + torso_angle) / 2
    pred = torch.cat((legs_pred, torso_pred), dim=1)
    pred = torch.cat((torch.zeros_like(pred[:, :1]), pred[:, 1:]), dim=1)  # pred[:,0] = 0
    R = _rotation(props, eps_x, u_y)
    p3 = _lift_centered(u, pred, depth)
    rot, rot_2d = _rotate_project(R, p3, depth)
    out = {}
    z, ld = F.inn_forward(G.part_2d(rot_2d, G.LEG_JOINTS), leg_flow)
    out["leg_likeli"] = F.nll(z, ld).mean()
    z, ld = F.inn_forward(G.part_2d(rot_2d, G.TORSO_JOINTS), torso_flow)
    out["torso_likeli"] = F.nll(z, ld).mean()
    out["likeli"] = out["torso_likeli"] + out["leg_likeli"]
    legs_rot, _ = N.lifter_forward(G.part_2d(rot_2d, G.LEG_JOINTS), leg, pose_only=True)
    torso_rot, _ = N.lifter_forward(G.part_2d(rot_2d, G.TORSO_JOINTS), torso, pose_only=True)
    pred_rot = torch.cat((legs_rot, torso_rot), dim=1)
    pred_rot = torch.cat((torch.zeros_like(pred_rot[:, :1]), pred_rot[:, 1:]), dim=1)
    out["L3d"], out["rep_rot"], out["re_rot_3d"], out["bl_prior"] = _consistency_terms(
        u, R, p3, rot, rot_2d, pred_rot, bone_rel, depth)
    out["loss"] = _total(out, cfg)
    if aux is not None:
        aux.update(pred=pred, props=props, R=R, pred_3d=p3, rot_poses=rot, rot_2d=rot_2d, pred_rot=pred_rot)
    return out


def lr_step(u, left, right, left_flow, right_flow, eps_x, u_y, cfg=DEFAULT_CFG, bone_rel=None, aux=None):
    """Left/right lifter training step, train_left_right_lifter.py:142-423 (after sampling)."""
    depth = cfg["depth"]
    if bone_rel is None:
        bone_rel = torch.tensor(G.BONE_REL_H36M, dtype=u.dtype)  # :76-79
    left_in, right_in = G.split_data_left_right(u)
    left_pred, left_angle = N.lifter_forward(left_in, left)
    right_pred, right_angle = N.lifter_forward(right_in, right)
    props = (left_angle + right_angle) / 2
    zero0 = lambda t: torch.cat((torch.zeros_like(t[:, :1]), t[:, 1:]), dim=1)
    pred_l = zero0(G.combine_left_right_1d(left_pred, right_pred, "left"))
    pred_r = zero0(G.combine_left_right_1d(left_pred, right_pred, "right"))
    R = _rotation(props, eps_x, u_y)
    p3_r = _lift_centered(u, pred_r, depth)
    p3_l = _lift_centered(u, pred_l, depth)
    rot_r, rot2d_r = _rotate_project(R, p3_r, depth)
    rot_l, rot2d_l = _rotate_project(R, p3_l, depth)
    norm_left_side, _ = G.split_data_left_right(rot2d_l)
    _, norm_right_side = G.split_data_left_right(rot2d_r)
    out = {}
    z, ld = F.inn_forward(norm_left_side, left_flow)
    out["likeli_right"] = F.nll(z, ld).mean()      # names swapped in the reference (:334-342)
    z, ld = F.inn_forward(norm_right_side, right_flow)
    out["likeli_left"] = F.nll(z, ld).mean()
    out["likeli"] = out["likeli_left"] + out["likeli_right"]
    pr_left, _ = N.lifter_forward(norm_left_side, left, pose_only=True)
    pr_right, _ = N.lifter_forward(norm_right_side, right, pose_only=True)
    prf_l = zero0(G.combine_left_right_1d(pr_left, pr_right, "left"))
    prf_r = zero0(G.combine_left_right_1d(pr_left, pr_right, "right"))
    tr = _consistency_terms(u, R, p3_r, rot_r, rot2d_r, prf_r, bone_rel, depth)
    tl = _consistency_terms(u, R, p3_l, rot_l, rot2d_l, prf_l, bone_rel, depth)
    for i, k in enumerate(("L3d", "rep_rot", "re_rot_3d", "bl_prior")):
        out[k] = tr[i] + tl[i]
    out["loss"] = _total(out, cfg)
    if aux is not None:
        aux.update(pred_left=pred_l, pred_right=pred_r, props=props, R=R, rot_2d_left=rot2d_l,
                   rot_2d_right=rot2d_r, pred_3d_left=p3_l, pred_3d_right=p3_r)
    return out


OCC_NAMES = ("left_arm", "right_arm", "left_leg", "right_leg", "left_side", "right_side", "both_legs", "torso")


def occ_targets_inputs(pose):
    """train_occlusion_models.py:176-191 on pose [B,3,17] -> (targets, inputs), dicts by OCC_NAMES."""
    t = {
        "left_arm": pose[:, :, 11:14].reshape(-1, 9),
        "right_arm": pose[:, :, 14:].reshape(-1, 9),
        "left_leg": pose[:, :, 4:7].reshape(-1, 9),
        "right_leg": pose[:, :, 1:4].reshape(-1, 9),
        "left_side": torch.cat((pose[:, :, 4:7], pose[:, :, 11:14]), dim=2).reshape(-1, 18),
        "right_side": torch.cat((pose[:, :, 1:4], pose[:, :, 14:]), dim=2).reshape(-1, 18),
        "both_legs": pose[:, :, 1:7].reshape(-1, 18),
        "torso": pose[:, :, 7:].reshape(-1, 30),
    }
    no_right_side, no_left_side = G.split_data_left_right_3d(pose)   # (left-list, right-list), :191
    i = {
        "left_arm": torch.cat((pose[:, :, :11], pose[:, :, 14:]), dim=2).reshape(-1, 42),
        "right_arm": pose[:, :, :14].reshape(-1, 42),
        "left_leg": torch.cat((pose[:, :, :4], pose[:, :, 7:]), dim=2).reshape(-1, 42),
        "right_leg": torch.cat((pose[:, :, :1], pose[:, :, 4:]), dim=2).reshape(-1, 42),
        "torso": pose[:, :, :7].reshape(-1, 21),
        "both_legs": torch.cat((pose[:, :, :1], pose[:, :, 7:]), dim=2).reshape(-1, 33),
        "left_side": no_left_side,
        "right_side": no_right_side,
    }
    return t, i


def occlusion_step(x, leg, torso, predictors, u_y1, u_y2, cfg=DEFAULT_CFG):
    """train_occlusion_models.py:150-302.  predictors: dict OCC_NAMES -> params.
    (The reference also runs the left/right lifters, :160-161; their outputs are dead.)"""
    depth = cfg["depth"]
    with torch.no_grad():  # lifters are frozen (:535-545); no grad reaches them
        legs_pred, _ = N.lifter_forward(G.part_2d(x, G.LEG_JOINTS), leg, pose_only=True)
        torso_pred, _ = N.lifter_forward(G.part_2d(x, G.TORSO_JOINTS), torso, pose_only=True)
        pred = torch.cat((legs_pred, torso_pred), dim=1)
        pred[:, 0] = 0.0
        pose = _lift_centered(x, pred, depth, clamp=False)     # no clamp here (:167)
    losses = {n: 0.0 for n in OCC_NAMES}
    for rnd in range(3):
        if rnd > 0:
            u = (u_y1, u_y2)[rnd - 1].reshape(-1, 1)
            zeros = torch.zeros_like(u)
            Ry = G.euler_angles_to_matrix(torch.cat((zeros, (u - 0.5) * 1.99 * np.pi, zeros), dim=1), "XYZ")
            pose = Ry.matmul(pose)
        tg, inp = occ_targets_inputs(pose)
        for n in OCC_NAMES:
            out = N.predictor_forward(inp[n], predictors[n])
            losses[n] = losses[n] + (out - tg[n]).square().sum(dim=1).mean()
    res = {"threed_loss_" + n: losses[n] for n in OCC_NAMES}
    res["loss"] = sum(losses[n] for n in OCC_NAMES)
    return res


def flow_step(x, flow_params, noise):
    """train_full_pose_norm_flow.py:75-93."""
    z, ld = F.inn_forward(x, flow_params)
    out = {"dist_2d": F.nll(z, ld).mean()}
    with torch.no_grad():
        zn = G.add_noise(z, noise, 0.2)
        s, _ = F.inn_forward(zn, flow_params, rev=True)
        s = s.reshape(-1, 2, 17).clone()
        s[:, :, [0]] = 0.0
        s = s.reshape(-1, x.shape[1])
    z2, ld2 = F.inn_forward(s, flow_params)
    out["dist_2d_sample"] = F.nll(z2, ld2).mean()
    out["loss"] = out["dist_2d"] + out["dist_2d_sample"]
    return out


def eval_lr_predict(poses_2d, left, right, choice="right", depth=10.0):
    """eval_h36m.py:50-78 -> predicted 3D poses [M,51] (not root-centred, no clamp)."""
    with torch.no_grad():
        inp_left, inp_right = G.split_data_left_right(poses_2d)
        pl, _ = N.lifter_forward(inp_left, left, pose_only=True)
        pr, _ = N.lifter_forward(inp_right, right, pose_only=True)
        pl = pl.clone(); pr = pr.clone()
        pl[:, 0] = 0.0
        pr[:, 0] = 0.0
        d = G.combine_left_right_1d(pl, pr, choice).reshape(-1, 17) + depth
        return G.lift(poses_2d, d).reshape(-1, 51)


def eval_lt_predict(poses_2d, leg, torso, depth=10.0):
    """train_leg_torso_lifter.py:295-309 -> [M,51]."""
    with torch.no_grad():
        a, _ = N.lifter_forward(G.part_2d(poses_2d, G.LEG_JOINTS), leg, pose_only=True)
        b, _ = N.lifter_forward(G.part_2d(poses_2d, G.TORSO_JOINTS), torso, pose_only=True)
        pred = torch.cat((a, b), dim=1)
        pred[:, 0] = 0.0
        return G.lift(poses_2d, pred + depth).reshape(-1, 51)


def eval_metrics(gt_3d, pred_3d, loop=False):
    """eval_h36m.py:83-97: PA-MPJPE (numpy 'best', fp64) and N-MPJPE (batch, scaled)."""
    gt_np = gt_3d.detach().cpu().numpy()
    pr_np = pred_3d.detach().cpu().numpy()
    if loop:  # the reference's literal per-pose loop
        pa = float(np.mean([M.pmpjpe_best_np(gt_np[i].reshape(-1, 51), pr_np[i].reshape(-1, 51), "best")
                            for i in range(gt_np.shape[0])]))
    else:
        pa = float(M.pmpjpe_best_batch(gt_np, pr_np).mean())
    n_mpjpe = float(M.mpjpe(gt_3d, pred_3d, num_joints=17, root_joint=0).mean())
    return {"pa_mpjpe": pa, "n_mpjpe": n_mpjpe}


def params_require_grad(p, flag=True):
    return {k: v.detach().clone().requires_grad_(flag) for k, v in p.items()}


def make_adam(param_dicts, lr=2e-4, weight_decay=1e-5):
    """train_leg_torso_lifter.py:111-114 -- one Adam per network."""
    return [torch.optim.Adam(list(p.values()), lr=lr, weight_decay=weight_decay) for p in param_dicts]


def part_flow_step(x, full_params, part_params, noise):
    """train_leg_torso_left_right_norm_flow.py:100-166.  part_params: dict legs/torso/left/right -> flow params.
    Returns the eight NLL means and their sum."""
    def parts(p):
        left, right = G.split_data_left_right(p)
        r = p.reshape(-1, 2, 17)
        return {"legs": r[:, :, :7].reshape(-1, 14), "torso": r[:, :, 7:].reshape(-1, 20), "left": left, "right": right}
    with torch.no_grad():
        z, _ = F.inn_forward(x, full_params)
        s, _ = F.inn_forward(G.add_noise(z, noise, 0.2), full_params, rev=True)
        s = s.reshape(-1, 2, 17).clone()
        s[:, :, [0]] = 0.0
        s = s.reshape(-1, x.shape[1])
    out = {}
    for tag, rows in (("", parts(x)), ("_sample", parts(s))):
        for n, inp in rows.items():
            zz, ld = F.inn_forward(inp, part_params[n])
            out["dist_2d_%s%s" % (n, tag)] = F.nll(zz, ld).mean()
    out["loss"] = sum(out.values())
    return out


def occ_validation_poses(poses_2d, legs_pred, torso_pred, left_pred, right_pred, predictors, depth=10.0):
    """train_occlusion_models.py:327-398 (validation_step up to the eight camera-frame poses), written out case by case
    like the script.  predictors: dict name -> callable([M,3k]) -> [M,3(17-k)].  Returns (inputs, full poses)."""
    left_split, right_split = G.split_data_left_right(poses_2d)
    legs_split = poses_2d.reshape(-1, 2, 17)[:, :, :7].reshape(-1, 14)
    torso_split = poses_2d.reshape(-1, 2, 17)[:, :, 7:].reshape(-1, 20)
    left_pred, right_pred = left_pred.clone(), right_pred.clone()
    left_pred[:, 0] = 0.0
    right_pred[:, 0] = 0.0
    left_pred, right_pred = left_pred + depth, right_pred + depth
    pred_lt = torch.cat((legs_pred, torso_pred), dim=1).clone()
    pred_lt[:, 0] = 0.0
    pred_lt = pred_lt + depth

    def lift(split, d, k):
        return torch.cat(((split.reshape(-1, 2, k) * d.reshape(-1, 1, k)).reshape(-1, 2 * k), d), dim=1).reshape(-1, 3, k)
    legs3, torso3 = lift(legs_split, pred_lt[:, :7], 7), lift(torso_split, pred_lt[:, 7:], 10)
    left3, right3 = lift(left_split, left_pred, 11), lift(right_split, right_pred, 11)
    torso3 = torso3 - legs3[:, :, [0]]
    legs3 = legs3 - legs3[:, :, [0]]
    left3 = left3 - left3[:, :, [0]]
    right3 = right3 - right3[:, :, [0]]
    inp = {"la": torch.cat((legs3, right3[:, :, 4:]), dim=2).reshape(-1, 42),
           "ra": torch.cat((legs3, left3[:, :, 4:]), dim=2).reshape(-1, 42),
           "ll": torch.cat((right3[:, :, :4], torso3), dim=2).reshape(-1, 42),
           "rl": torch.cat((left3[:, :, :4], torso3), dim=2).reshape(-1, 42),
           "torso": legs3.reshape(-1, 21),
           "legs": torch.cat((legs3[:, :, [0]], torso3), dim=2).reshape(-1, 33),
           "right": left3.reshape(-1, 33),           # no_right_side
           "left": right3.reshape(-1, 33)}           # no_left_side
    name = {"la": "left_arm", "ra": "right_arm", "ll": "left_leg", "rl": "right_leg", "torso": "torso",
            "legs": "both_legs", "left": "left_side", "right": "right_side"}
    out = {k: predictors[name[k]](v) for k, v in inp.items()}

    def limb(pose, l, which):                         # combine_pose_and_limb, :67-78
        l, pose = l.reshape(-1, 3, 3), pose.reshape(-1, 3, 14)
        cut = {"ll": 4, "rl": 1, "la": 11, "ra": 14}[which]
        return torch.cat((pose[:, :, :cut], l, pose[:, :, cut:]), dim=2).reshape(-1, 51)
    full = {k: limb(inp[k], out[k], k) for k in ("la", "ra", "ll", "rl")}
    full["torso"] = torch.cat((inp["torso"].reshape(-1, 3, 7), out["torso"].reshape(-1, 3, 10)), dim=2).reshape(-1, 51)
    il = inp["legs"].reshape(-1, 3, 11)
    full["legs"] = torch.cat((il[:, :, :1], out["legs"].reshape(-1, 3, 6), il[:, :, 1:]), dim=2).reshape(-1, 51)
    for side in ("left", "right"):                    # utils/helpers.py:121-136
        o, v = out[side].reshape(-1, 3, 6), inp[side].reshape(-1, 3, 11)
        if side == "right":
            cols = [v[:, :, 0], o[:, :, 0], o[:, :, 1], o[:, :, 2]] + [v[:, :, i] for i in range(1, 11)] + \
                   [o[:, :, 3], o[:, :, 4], o[:, :, 5]]
        else:
            cols = [v[:, :, 0], v[:, :, 1], v[:, :, 2], v[:, :, 3], o[:, :, 0], o[:, :, 1], o[:, :, 2], v[:, :, 4], v[:, :, 5],
                    v[:, :, 6], v[:, :, 7], o[:, :, 3], o[:, :, 4], o[:, :, 5], v[:, :, 8], v[:, :, 9], v[:, :, 10]]
        full[side] = torch.stack(cols, dim=2).reshape(-1, 51)
    glob = {k: torch.cat((p[:, :34], p[:, 34:51] + depth), dim=1) for k, p in full.items()}
    return inp, glob
