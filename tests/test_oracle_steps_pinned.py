"""CPU: the step restatements in oracle/steps.py against the reference's OWN training_step code.

tests/golden/ref_steps.npz was produced by oracle/gen_golden.py::reference_training_steps: each script's training_step
method is cut out of its AST and executed in the build container on a stand-in `self` that carries the reference's
networks and -- the one substitution, FrEIA being unavailable -- the oracle flow.  Here the same seeded inputs and the
same random draws (re-drawn in the order the reference code consumes them) go through oracle.steps; losses must agree
to fp32 round-off and the weight gradients the reference's autograd produced must match."""
import numpy as np
import pytest
import torch

from oracle import flow as OF, nets as ON, steps as OS

B = 16


def _close(a, b, rtol=2e-5):
    assert abs(a - b) <= rtol * abs(b) + 1e-7, (a, b)


def _grad_close(got, ref, rtol=2e-4):
    got, ref = got.detach().numpy(), ref
    assert np.abs(got - ref).max() <= rtol * np.abs(ref).max() + 1e-8, np.abs(got - ref).max() / np.abs(ref).max()


@pytest.fixture(scope="module")
def G(golden_module):
    return golden_module["ref_steps"]


@pytest.fixture(scope="module")
def golden_module():
    import os
    d = os.path.join(os.path.dirname(__file__), "golden")
    return {n[:-4]: np.load(os.path.join(d, n)) for n in os.listdir(d) if n.endswith(".npz")}


def _draws(seed, with_sampling=True):
    torch.manual_seed(seed)
    noise = torch.randn(B, 34)                                                  # add_noise: randn_like(z) (helpers.py:300)
    eps = torch.normal(torch.zeros((2 * B, 1)), torch.ones((2 * B, 1))).reshape(-1)   # elevation draw (:169-171)
    u_y = torch.rand((2 * B, 1)).reshape(-1)                                    # azimuth draw (:176)
    return noise, eps, u_y


def test_leg_torso_step_equals_reference_training_step(G):
    x = torch.from_numpy(G["x"])
    full = OF.init_flow_params(34, 40, perturb=0.3)
    leg, torso = OS.params_require_grad(ON.init_lifter_params(7, 11)), OS.params_require_grad(ON.init_lifter_params(10, 12))
    noise, eps, u_y = _draws(1001)
    u = OS.sample_poses(x, full, noise)
    out = OS.lt_step(u, leg, torso, OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3), eps, u_y)
    for k in ("leg_likeli", "torso_likeli", "likeli", "L3d", "rep_rot", "re_rot_3d", "bl_prior", "loss"):
        _close(out[k].item(), float(G["lt_" + k]))
    out["loss"].backward()
    _grad_close(leg["upscale.weight"].grad, G["lt_dW_leg_upscale"])
    _grad_close(torso["angles.weight"].grad, G["lt_dW_torso_angles"])


def test_left_right_step_equals_reference_training_step(G):
    x = torch.from_numpy(G["x"])
    full = OF.init_flow_params(34, 40, perturb=0.3)
    left, right = OS.params_require_grad(ON.init_lifter_params(11, 13)), OS.params_require_grad(ON.init_lifter_params(11, 14))
    noise, eps, u_y = _draws(1002)
    u = OS.sample_poses(x, full, noise)
    out = OS.lr_step(u, left, right, OF.init_flow_params(22, 43, perturb=0.3), OF.init_flow_params(22, 44, perturb=0.3), eps, u_y)
    for k in ("likeli_right", "likeli_left", "likeli", "L3d", "rep_rot", "re_rot_3d", "bl_prior", "loss"):
        _close(out[k].item(), float(G["lr_" + k]))
    out["loss"].backward()
    _grad_close(left["upscale.weight"].grad, G["lr_dW_left_upscale"])
    _grad_close(right["downscale.weight"].grad, G["lr_dW_right_downscale"])


def test_occlusion_step_equals_reference_training_step(G):
    x = torch.from_numpy(G["x"])
    leg, torso = ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)
    nin = {"left_arm": 14, "right_arm": 14, "left_leg": 14, "right_leg": 14, "left_side": 11, "right_side": 11,
           "both_legs": 11, "torso": 7}
    nout = {"left_arm": 9, "right_arm": 9, "left_leg": 9, "right_leg": 9, "left_side": 18, "right_side": 18,
            "both_legs": 18, "torso": 30}
    preds = {n: OS.params_require_grad(ON.init_predictor_params(nin[n], nout[n], 100 + i)) for i, n in enumerate(OS.OCC_NAMES)}
    torch.manual_seed(1003)
    u1, u2 = torch.rand((B, 1)).reshape(-1), torch.rand((B, 1)).reshape(-1)     # Ry draws (:213, :256)
    out = OS.occlusion_step(x, leg, torso, preds, u1, u2)
    for n in OS.OCC_NAMES:
        _close(out["threed_loss_" + n].item(), float(G["occ_threed_loss_" + n]))
    _close(out["loss"].item(), float(G["occ_loss"]))
    out["loss"].backward()
    _grad_close(preds["torso"]["downscale.weight"].grad, G["occ_dW_torso_downscale"])
    _grad_close(preds["left_side"]["upscale.weight"].grad, G["occ_dW_left_upscale"])


def test_leg_torso_validation_equals_reference_validation_step(G):
    """train_leg_torso_lifter.py:286-337: PA-MPJPE ('best', per-pose numpy loop in the reference), scale-matched MPJPE,
    AUC and PCK of the lifted validation poses."""
    from oracle import geometry as OG, metrics as OM
    x, gt = torch.from_numpy(G["val_x"]), torch.from_numpy(G["val_gt"])
    leg, torso = ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)
    with torch.no_grad():
        lp, _ = ON.lifter_forward(OG.part_2d(x, OG.LEG_JOINTS), leg)
        tp, _ = ON.lifter_forward(OG.part_2d(x, OG.TORSO_JOINTS), torso)
        pred = torch.cat((lp, tp), dim=1)
        pred[:, 0] = 0.0
        poses = OG.lift(x, pred + 10.0).reshape(-1, 51)
    _close(float(OM.pmpjpe_best_batch(gt.numpy(), poses.numpy()).mean()), float(G["ltval_pa"]), rtol=1e-5)
    _close(OM.mpjpe(gt, poses, num_joints=17, root_joint=0).mean().item(), float(G["ltval_mpjpe_scaled"]), rtol=1e-5)
    _close(float(OM.auc(gt, poses, num_joints=17, root_joint=0)), float(G["ltval_auc"]), rtol=1e-5)
    _close(float(OM.pck(gt, poses, num_joints=17, root_joint=0)), float(G["ltval_pck"]), rtol=1e-6)


def test_occlusion_validator_equals_reference_validation_step(G):
    """train_occlusion_models.py:317-509 executed by the reference's own code vs links_b200.occ_assembly's
    OcclusionValidator (the product's assembly logic) wired to the oracle networks and metrics."""
    from links_b200 import occ_assembly as OA
    from test_occ_assembly_cpu import _OracleMetrics
    lp = {"legs": ON.init_lifter_params(7, 11), "torso": ON.init_lifter_params(10, 12),
          "left": ON.init_lifter_params(11, 13), "right": ON.init_lifter_params(11, 14)}
    pp = {n: ON.init_predictor_params(len(OA.visible_joints(c)), 3 * len(OA.CASES[c][2]), 100 + OS.OCC_NAMES.index(n))
          for c, (n, _, _) in OA.CASES.items()}
    val = OA.OcclusionValidator({k: (lambda v, p=p: ON.lifter_forward(v, p)) for k, p in lp.items()},
                                {n: (lambda v, p=p: ON.predictor_forward(v, p)) for n, p in pp.items()}, _OracleMetrics(), 10.0)
    got = val.run(torch.from_numpy(G["val_x"]), torch.from_numpy(G["val_gt"]))
    for c in OA.CASES:
        _close(got["pa_" + c], float(G["occval_pa_" + c]), rtol=2e-5)
        _close(got["mpjpe_scaled_" + c], float(G["occval_mpjpe_scaled_" + c]), rtol=2e-5)


def test_left_right_validation_equals_reference_validation_step(G):
    """train_left_right_lifter.py:437-511 (the lift eval_h36m.py:50-97 also performs): both combine choices."""
    from oracle import metrics as OM
    x, gt = torch.from_numpy(G["val_x"]), torch.from_numpy(G["val_gt"])
    left, right = ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)
    for choice in ("left", "right"):
        poses = OS.eval_lr_predict(x, left, right, choice=choice, depth=10.0)
        _close(float(OM.pmpjpe_best_batch(gt.numpy(), poses.numpy()).mean()), float(G["lrval_pa_" + choice]), rtol=1e-5)
        _close(OM.mpjpe(gt, poses, num_joints=17, root_joint=0).mean().item(), float(G["lrval_mpjpe_scaled_" + choice]),
               rtol=1e-5)


def test_eval_restatement_equals_reference_eval_script(G):
    """eval_h36m.py:46-97 (the script's own top-level statements, executed by gen_golden.py) vs
    oracle.steps.eval_lr_predict + eval_metrics -- the pair every GPU eval parity test is checked against."""
    x, gt = torch.from_numpy(G["val_x"]), torch.from_numpy(G["val_gt"])
    left, right = ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)
    poses = OS.eval_lr_predict(x, left, right, choice="right", depth=10.0)
    for loop in (False, True):
        out = OS.eval_metrics(gt, poses, loop=loop)
        _close(out["pa_mpjpe"], float(G["evalh36m_pa"]), rtol=1e-5)
        _close(out["n_mpjpe"], float(G["evalh36m_mpjpe_scaled"]), rtol=1e-5)


def test_flow_trainers_equal_reference_loop_bodies(G):
    """train_full_pose_norm_flow.py:67-98 and train_leg_torso_left_right_norm_flow.py:100-174: the bodies of the scripts'
    batch loops (executed by gen_golden.py with the oracle flow standing in for FrEIA) vs oracle.steps.flow_step /
    part_flow_step -- the references of the GPU FlowTrainStep / PartFlowTrainer tests."""
    x = torch.from_numpy(G["x"])
    fp = OS.params_require_grad(OF.init_flow_params(34, 45, perturb=0.3))
    torch.manual_seed(1004)
    out = OS.flow_step(x, fp, torch.randn(B, 34))
    for k in ("dist_2d", "dist_2d_sample", "loss"):
        _close(out[k].item(), float(G["flowtrain_" + k]))
    out["loss"].backward()
    _grad_close(fp["module_list.2.subnet.2.weight"].grad, G["flowtrain_dW"])
    full = OF.init_flow_params(34, 40, perturb=0.3)
    parts = {n: OS.params_require_grad(OF.init_flow_params(w, 60 + i, perturb=0.3))
             for i, (n, w) in enumerate((("legs", 14), ("torso", 20), ("left", 22), ("right", 22)))}
    torch.manual_seed(1005)
    out = OS.part_flow_step(x, full, parts, torch.randn(B, 34))
    for k in out:
        _close(out[k].item(), float(G["partflow_" + k]), rtol=5e-5)
    out["loss"].backward()
    _grad_close(parts["left"]["module_list.1.subnet.0.weight"].grad, G["partflow_dW_left"])
