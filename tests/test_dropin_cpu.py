"""CPU: host-side drop-in helpers (utils/helpers.py, utils/rotation_conversions.py, utils/metrics.py) against the
golden vectors produced by the REFERENCE's own modules (tests/golden/*.npz, oracle/gen_golden.py), the oracle
pinned to those vectors, and the drop-in modules' state-dict key contract."""
import numpy as np
import pytest
import torch

from oracle import geometry as OG
from oracle import metrics as OM
from utils import helpers as H
from utils import metrics as MN
from utils import rotation_conversions as RC


def test_oracle_and_helpers_match_reference_index_maps(golden):
    G = golden["index_maps"]
    a34 = torch.from_numpy(G["split_lr_in"])
    for mod in (H, OG):
        l, r = mod.split_data_left_right(a34)
        assert np.array_equal(l.numpy(), G["split_lr_left"]) and np.array_equal(r.numpy(), G["split_lr_right"])
        l3, r3 = mod.split_data_left_right_3d(torch.from_numpy(G["split3d_in"]))
        assert np.array_equal(l3.numpy(), G["split3d_left"]) and np.array_equal(r3.numpy(), G["split3d_right"])
    la, ra = torch.from_numpy(G["combine_la"]), torch.from_numpy(G["combine_ra"])
    for ch in ("left", "right"):
        assert np.array_equal(H.combine_left_right_pred_1d(la, ra, ch).reshape(-1, 17).numpy(), G["combine1d_" + ch])
        assert np.array_equal(OG.combine_left_right_1d(la, ra, ch).numpy(), G["combine1d_" + ch])
        for dims, fn in ((2, H.combine_left_right_pred_2d), (3, H.combine_left_right_pred_3d)):
            ll = torch.arange(6 * dims * 11, dtype=torch.float32).reshape(6, dims * 11)
            assert np.array_equal(fn(ll, 5000 + ll, ch).numpy(), G["combine%dd_%s" % (dims, ch)])


def test_oracle_and_helpers_match_reference_geometry(golden):
    G = golden["geometry"]
    ang = torch.from_numpy(G["euler_in"])
    for conv in ("XYZ", "ZYX", "YXZ", "XZY"):
        assert np.array_equal(OG.euler_angles_to_matrix(ang, conv).numpy(), G["euler_" + conv])
        np.testing.assert_allclose(RC.euler_angles_to_matrix(ang, conv).numpy(), G["euler_" + conv], atol=1e-6)
    p3 = torch.from_numpy(G["proj_in"])
    for mod in (H, OG):
        assert np.array_equal(mod.perspective_projection(p3).numpy(), G["proj_out"])
        assert np.array_equal(mod.get_bone_lengths_all(p3).numpy(), G["bones_out"])
        assert np.array_equal(mod.normalize_head(G["normhead_in"].copy()), G["normhead_out"])
        assert np.array_equal(mod.normalize_head_test(G["normhead_in"].copy()), G["normhead_test_out"])


def test_rotation_error_behaviour():
    """ValueError on malformed input, like rotation_conversions.py:51-59."""
    for fn in (RC.euler_angles_to_matrix, OG.euler_angles_to_matrix):
        with pytest.raises(ValueError):
            fn(torch.zeros(4, 2), "XYZ")
        with pytest.raises(ValueError):
            fn(torch.zeros(4, 3), "XY")
        with pytest.raises(ValueError):
            fn(torch.zeros(4, 3), "XXY")
        with pytest.raises(ValueError):
            fn(torch.zeros(4, 3), "XYW")


def test_oracle_and_numpy_metrics_match_reference(golden):
    G = golden["metrics"]
    gt, pred = G["gt"].astype(np.float64), G["pred"].astype(np.float64)
    m = MN.Metrics()
    for refl in ("best", True, False):
        ref = G["pmpjpe_np_%s" % refl]
        got = np.array([m.pmpjpe(gt[i].reshape(-1, 51), pred[i].reshape(-1, 51), reflection=refl) for i in range(16)])
        np.testing.assert_allclose(got, ref[:16], rtol=0, atol=1e-9)
        ora = np.array([OM.pmpjpe_best_np(gt[i].reshape(-1, 51), pred[i].reshape(-1, 51), reflection=refl) for i in range(16)])
        np.testing.assert_allclose(ora, ref[:16], rtol=0, atol=1e-9)
    np.testing.assert_allclose(OM.pmpjpe_best_batch(gt, pred), G["pmpjpe_np_best"], rtol=0, atol=1e-8)
    got = np.array([m.mpjpe(gt[i].reshape(-1, 51), pred[i].reshape(-1, 51)) for i in range(8)])
    np.testing.assert_allclose(got, G["mpjpe_np"], rtol=0, atol=1e-9)
    # torch metrics oracle vs reference outputs
    g, p = torch.from_numpy(G["gt"]), torch.from_numpy(G["pred"])
    assert np.array_equal(OM.mpjpe(g, p, num_joints=17, root_joint=0).numpy(), G["mpjpe_j17_s1"])
    np.testing.assert_allclose(OM.pmpjpe_batch(g, p, num_joints=17).numpy(), G["pmpjpe_batch_j17"], rtol=1e-4, atol=1e-3)


def test_oracle_networks_match_reference_outputs(golden):
    """oracle.nets reproduces the reference modules' outputs bit-exactly (weights regenerated from the seed)."""
    from oracle import nets as ON
    G = golden["nets"]
    for cls, nj in (("Leg_Lifter", 7), ("Torso_Lifter", 10), ("Left_Right_Lifter", 11)):
        p = ON.init_lifter_params(nj, int(G[cls + "_seed"]))
        xd, xa = ON.lifter_forward(torch.from_numpy(G[cls + "_x"]), p)
        assert np.array_equal(xd.numpy(), G[cls + "_xd"]) and np.array_equal(xa.numpy(), G[cls + "_xa"])
    p = ON.init_predictor_params(14, 9, int(G["Occluded_Limb_Predictor_seed"]))
    y = ON.predictor_forward(torch.from_numpy(G["Occluded_Limb_Predictor_x"]), p)
    assert np.array_equal(y.numpy(), G["Occluded_Limb_Predictor_y"])


def test_flow_oracle_known_answers():
    """FrEIA parity is unpinned; anchor the restatement with self-consistency known answers (fp64):
    invertibility, log_jac_det == slogdet of the autograd Jacobian, identity-at-init."""
    from oracle import flow as OF
    C = 14
    p = {k: v.double() for k, v in OF.init_flow_params(C, 3, perturb=0.3).items()}
    x = torch.randn(5, C, dtype=torch.float64) * 0.3
    z, ld = OF.inn_forward(x, p)
    xr, ldr = OF.inn_forward(z, p, rev=True)
    # w_perm_inv is the fp32-rounded transpose of w_perm, so the inverse is exact to ~1e-7 per block
    assert (xr - x).abs().max() < 1e-5 and (ld + ldr).abs().max() < 1e-5
    for i in range(2):
        Jm = torch.autograd.functional.jacobian(lambda v: OF.inn_forward(v[None], p)[0][0], x[i])
        # w_perm is an fp32-rounded orthogonal matrix (|log det| ~ 1e-8 per block), which FrEIA does not count
        assert abs(torch.linalg.slogdet(Jm)[1].item() - ld[i].item()) < 1e-6
    p0 = {k: v.double() for k, v in OF.init_flow_params(C, 3, perturb=0.0).items()}
    g = 0.1 * torch.nn.functional.softplus(p0["module_list.0.global_scale"], beta=0.5)
    assert (g - 1).abs().max() < 1e-6            # global affine starts as the identity (init value stored in fp32)


def test_dropin_state_dict_contract():
    """Key names / counts of the drop-in modules equal the reference's: lifters 62 keys incl. the
    unused LayerNorms, predictors 36 incl. unused res_common, flows 8 keys per block."""
    from utils.models_def import Leg_Lifter, Occluded_Limb_Predictor
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    sd = Leg_Lifter(use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.25).state_dict()
    assert len(sd) == 62 and "res_angle2.bn1.weight" in sd and sd["upscale.weight"].shape == (1024, 14)
    assert sd["downscale.weight"].shape == (7, 1024) and sd["angles.bias"].shape == (1,)
    sd = Occluded_Limb_Predictor(use_batchnorm=False, num_joints=14).state_dict()
    assert len(sd) == 36 and "res_common.l1.weight" in sd and sd["downscale.weight"].shape == (9, 1024)
    inn = Ff.SequenceINN(34)
    for _ in range(8):
        inn.append(Fm.AllInOneBlock, subnet_constructor=H.subnet_fc, permute_soft=True)
    sd = inn.state_dict()
    assert len(sd) == 64
    assert sd["module_list.3.subnet.0.weight"].shape == (1024, 17) and sd["module_list.3.subnet.2.weight"].shape == (34, 1024)
    assert sd["module_list.0.global_scale"].shape == (1, 34) and sd["module_list.7.w_perm_inv"].shape == (34, 34)
    w = sd["module_list.0.w_perm"]
    assert torch.allclose(w @ w.t(), torch.eye(34), atol=1e-5)
    with pytest.raises(NotImplementedError):
        Leg_Lifter(use_batchnorm=True)


def test_remaining_helpers_match_reference(golden):
    """utils/helpers.py functions no step body calls (alternative splits, occluded-pose assembly, part bone lengths,
    fixed-scale normalisers, part projections, latent interpolation, occlusion masks) against outputs of the reference
    itself (tests/golden/helpers_extra.npz, oracle/gen_golden.py::helpers_extra)."""
    import random
    G = golden["helpers_extra"]
    a34 = torch.from_numpy(G["a34"])
    l, r = H.split_data_left_right_v2(a34)
    assert np.array_equal(l.numpy(), G["split_v2_left"]) and np.array_equal(r.numpy(), G["split_v2_right"])
    ln, rn = H.split_data_left_right_numpy(G["a34"])
    assert np.array_equal(ln, G["split_np_left"]) and np.array_equal(rn, G["split_np_right"])
    l, r = H.temporal_split_data_left_right(torch.from_numpy(G["a68"]))
    assert np.array_equal(l.numpy(), G["split_temporal_left"]) and np.array_equal(r.numpy(), G["split_temporal_right"])
    occ, vis = torch.from_numpy(G["occ_part"]), torch.from_numpy(G["vis_part"])
    for side in ("right", "left"):
        assert np.array_equal(H.combine_left_right_occluded_3d(occ, vis, side).numpy(), G["combine_occluded_" + side])
    np.testing.assert_allclose(H.get_bone_lengths_legs(torch.from_numpy(G["legs3d"])).numpy(), G["bones_legs"], rtol=1e-6)
    np.testing.assert_allclose(H.get_bone_lengths_left_right(torch.from_numpy(G["lr3d"])).numpy(), G["bones_left_right"],
                               rtol=1e-6)
    for fn in ("normalize_head_test_mpi_chest", "normalize_head_test_mpi_vnect", "normalize_head_test_temporal"):
        np.testing.assert_allclose(getattr(H, fn)(G["raw2d"].copy()), G[fn], rtol=1e-14)
    np.testing.assert_array_equal(H.interpolate_gaussian_batch(torch.from_numpy(G["latent"]), 0.3).numpy(), G["interp_0.3"])
    for fn in ("perspective_projection_legs", "perspective_projection_torso", "perspective_projection_left_right"):
        np.testing.assert_array_equal(getattr(H, fn)(torch.from_numpy(G[fn + "_in"])).numpy(), G[fn + "_out"])
    random.seed(77)
    np.testing.assert_array_equal(H.occlusion_create(a34).numpy(), G["occlusion_create"])


def test_module_contract_matches_reference():
    """Every class of the reference's utils/models_def.py: constructor parameter names / defaults and state_dict key
    names / order / shapes of the drop-in equal the reference's (tests/golden/module_contract.json, written by
    oracle/gen_golden.py::module_contract from the reference itself)."""
    import inspect
    import json
    import os
    from utils import models_def as MD
    with open(os.path.join(os.path.dirname(__file__), "golden", "module_contract.json")) as f:
        G = json.load(f)
    assert len(G) == 10
    for name, spec in G.items():
        cls = getattr(MD, name)
        sig = [[p.name, None if p.default is inspect._empty else p.default]
               for p in list(inspect.signature(cls.__init__).parameters.values())[1:]]
        assert sig == spec["signature"], (name, sig, spec["signature"])
        sd = cls(use_batchnorm=False, **spec["kwargs"]).state_dict()
        got = [[k, list(v.shape)] for k, v in sd.items()]
        assert got == spec["state_dict"], (name, [a for a, b in zip(got, spec["state_dict"]) if a != b][:3])


def test_numpy_metrics_full_surface_matches_reference(golden):
    """utils/metrics.py drop-in: every flag combination of mpjpe, PCK, and the complete (d, Z, tform) output of
    procrustes for all scaling / reflection settings, against the reference's outputs (metrics_np_extra.npz)."""
    import warnings
    from utils.metrics import Metrics
    G = golden["metrics_np_extra"]
    m = Metrics()
    gt, pred = G["gt"], G["pred"]
    for sc in (False, True):
        for ma in (False, True):
            got = np.array([m.mpjpe(gt[i:i + 1], pred[i:i + 1], scale=sc, mean_align=ma) for i in range(6)])
            np.testing.assert_allclose(got, G["mpjpe_s%d_m%d" % (sc, ma)], rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for refl in (False, True, "best"):
            got = np.array([m.PCK(gt[i:i + 1], pred[i:i + 1], reflection=refl) for i in range(6)])
            np.testing.assert_allclose(got, G["pck_%s" % refl], rtol=1e-10)
            for scaling in (True, False):
                d, Z, tf = m.procrustes(gt[0].reshape(3, 17).T, pred[0].reshape(3, 17).T, scaling=scaling, reflection=refl)
                tag = "proc_%s_%d_" % (refl, scaling)
                np.testing.assert_allclose(d, G[tag + "d"], rtol=1e-9)
                np.testing.assert_allclose(Z, G[tag + "Z"], rtol=1e-9, atol=1e-9)
                np.testing.assert_allclose(tf["rotation"], G[tag + "rot"], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(tf["scale"], G[tag + "scale"], rtol=1e-9)
                np.testing.assert_allclose(tf["translation"], G[tag + "trans"], rtol=1e-9, atol=1e-9)
