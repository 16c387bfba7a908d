"""CPU: occlusion inference assembly (reference train_occlusion_models.py:316-398).  The table-driven product code
(links_b200/occ_assembly.py) must equal (i) the reference's own combine_pose_and_limb (golden cut out of the script)
and (ii) the case-by-case restatement of the validation step in oracle/steps.py, bit for bit."""
import numpy as np
import pytest
import torch

from links_b200 import occ_assembly as OA
from oracle import steps as OS


def test_combine_pose_and_limb_matches_reference(golden):
    G = golden["script_funcs"]
    pose, limb = torch.from_numpy(G["pose"]), torch.from_numpy(G["limb"])
    for which in ("ll", "rl", "la", "ra"):
        np.testing.assert_array_equal(OA.combine_pose_and_limb(pose, limb, which).numpy(), G["combine_" + which])
    with pytest.raises(ValueError):
        OA.combine_pose_and_limb(pose, limb, "xx")


def test_case_table_is_consistent():
    from links_b200 import maps
    for case, (name, pieces, pred_joints) in OA.CASES.items():
        assert sorted(OA.visible_joints(case) + pred_joints) == list(range(17))
        assert pred_joints == maps.OCC_TARGET_JOINTS[name]                    # what the predictor was trained to output
        assert len(OA.visible_joints(case)) * 3 == len(maps.occ_input_index(name)[0]) // maps.occ_input_index(name)[1]


@pytest.mark.parametrize("M", [1, 6])
def test_assembly_matches_validation_restatement(M):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, 34, generator=g) * 0.1
    lp, tp = torch.randn(M, 7, generator=g), torch.randn(M, 10, generator=g)
    le, ri = torch.randn(M, 11, generator=g), torch.randn(M, 11, generator=g)
    W = {n: torch.randn(3 * len(OA.visible_joints(c)), 3 * len(OA.CASES[c][2]), generator=g) for c, (n, _, _) in OA.CASES.items()}
    preds = {n: (lambda v, n=n: torch.tanh(v @ W[n])) for n in W}
    inp_ref, glob_ref = OS.occ_validation_poses(x, lp, tp, le, ri, preds, 10.0)
    parts = OA.lift_parts(x, lp, tp, le, ri, 10.0)
    for c, (n, _, _) in OA.CASES.items():
        v = OA.visible_input(parts, c)
        assert torch.equal(v, inp_ref[c]), c
        assert torch.equal(OA.to_global(OA.assemble(c, v, preds[n](v)), 10.0), glob_ref[c]), c


class _OracleMetrics:
    """CPU stand-in with the interface of utils.metrics_batch.Metrics (the GPU drop-in)."""

    def pmpjpe_best(self, gt, pred):
        from oracle import metrics as OM
        return torch.from_numpy(OM.pmpjpe_best_batch(gt.numpy(), pred.numpy()))

    def mpjpe(self, gt, pred, num_joints=17, root_joint=0):
        from oracle import metrics as OM
        return OM.mpjpe(gt, pred, num_joints=num_joints, root_joint=root_joint)


def oracle_validator(seed=0):
    """OcclusionValidator wired to the oracle networks / metrics: what the GPU validator must reproduce."""
    from links_b200.synth import synth_poses
    from oracle import nets as ON
    lp = {"legs": ON.init_lifter_params(7, 11 + seed), "torso": ON.init_lifter_params(10, 12 + seed),
          "left": ON.init_lifter_params(11, 13 + seed), "right": ON.init_lifter_params(11, 14 + seed)}
    pp = {n: ON.init_predictor_params(len(OA.visible_joints(c)), 3 * len(OA.CASES[c][2]), 100 + i + seed)
          for i, (c, (n, _, _)) in enumerate(OA.CASES.items())}
    lifters = {k: (lambda x, p=p: ON.lifter_forward(x, p)) for k, p in lp.items()}
    predictors = {n: (lambda x, p=p: ON.predictor_forward(x, p)) for n, p in pp.items()}
    return OA.OcclusionValidator(lifters, predictors, _OracleMetrics(), 10.0), lp, pp, lifters, predictors


def test_validator_equals_step_by_step_validation():
    from links_b200.synth import synth_poses
    from oracle import metrics as OM
    val, lp, pp, lifters, predictors = oracle_validator()
    x2d, gt = synth_poses(24, seed=31)
    x, g = torch.from_numpy(x2d), torch.from_numpy(gt)
    got = val.run(x, g)
    # the same thing the long way round (oracle.steps.occ_validation_poses follows the script line by line)
    xx = x.reshape(-1, 2, 17)
    d = {k: lifters[k](xx[:, :, j].reshape(24, -1))[0] for k, j in OA.PART_JOINTS.items()}
    _, glob = OS.occ_validation_poses(x, d["legs"], d["torso"], d["left"], d["right"], predictors, 10.0)
    assert set(got) == {p + c for c in OA.CASES for p in ("pa_", "mpjpe_scaled_")}
    for c in OA.CASES:
        pa = OM.pmpjpe_best_batch(gt, glob[c].detach().numpy()).mean()
        mp = OM.mpjpe(g, glob[c].detach(), num_joints=17, root_joint=0).double().mean().item()
        assert abs(got["pa_" + c] - pa) < 1e-9 and abs(got["mpjpe_scaled_" + c] - mp) < 1e-9, c
