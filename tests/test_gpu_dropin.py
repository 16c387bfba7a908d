"""GPU: the drop-in module surface (utils.models_def, FrEIA shim, utils.metrics_batch) vs the oracle / golden vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_lifter_module_autograd(golden):
    from oracle import nets as ON
    from utils.models_def import Leg_Lifter
    G = golden["nets"]
    p = ON.init_lifter_params(7, int(G["Leg_Lifter_seed"]))
    m = Leg_Lifter(use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.25).cuda()
    missing = m.load_state_dict(p, strict=False)
    assert all(".bn" in k for k in missing.missing_keys)
    x = torch.from_numpy(G["Leg_Lifter_x"]).cuda().requires_grad_(True)
    xd, xa = m(x)
    # forward vs the REFERENCE module's output (fp32): depth offsets within 5e-3 abs (= 5e-4 relative on joints)
    assert (xd.detach().cpu() - torch.from_numpy(G["Leg_Lifter_xd"])).abs().max() < 5e-3
    assert (xa.detach().cpu() - torch.from_numpy(G["Leg_Lifter_xa"])).abs().max() < 5e-3
    (xd.square().sum() + xa.sum()).backward()
    assert rel_fro(x.grad.cpu(), torch.from_numpy(G["Leg_Lifter_dx"])) < 0.1
    gw = m.res_pose2.l1.weight.grad
    assert gw is not None and abs(gw.double().sum().item() - float(G["Leg_Lifter_dW_res_pose2_l1_sum"])) < 0.05 * float(G["Leg_Lifter_dW_res_pose2_l1_abs"])
    assert m.res_pose1.bn1.weight.grad is None          # unused LayerNorm never receives a gradient
    # an optimiser step through torch.optim works on the module's own parameters and is picked up next forward
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.step()
    xd2, _ = m(x.detach())
    assert (xd2 - xd.detach()).abs().max() > 0
    with torch.no_grad():
        xd3, _ = m(x.detach())
    assert torch.allclose(xd3, xd2)
    with pytest.raises(Exception):
        m(x.detach().cpu())


def test_predictor_module(golden):
    from oracle import nets as ON
    from utils.models_def import Occluded_Torso_Predictor
    G = golden["nets"]
    p = ON.init_predictor_params(7, 30, int(G["Occluded_Torso_Predictor_seed"]))
    m = Occluded_Torso_Predictor(use_batchnorm=False, num_joints=7).cuda()
    m.load_state_dict(p, strict=False)
    y = m(torch.from_numpy(G["Occluded_Torso_Predictor_x"]).cuda())
    assert rel_fro(y.detach().cpu(), torch.from_numpy(G["Occluded_Torso_Predictor_y"])) < 1.5e-2


def test_freia_shim_matches_oracle_and_fixture(golden):
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    from oracle import flow as OF
    from utils.helpers import subnet_fc
    G = golden["steps"]
    inn = Ff.SequenceINN(34)
    for _ in range(8):
        inn.append(Fm.AllInOneBlock, subnet_constructor=subnet_fc, permute_soft=True)
    params = OF.init_flow_params(34, 40, perturb=0.3)
    inn.load_state_dict(params)
    inn.cuda()
    for q in inn.parameters():
        q.requires_grad = False
    x = torch.from_numpy(G["x"]).cuda().requires_grad_(True)
    z, ld = inn(x)
    np.testing.assert_allclose(z.detach().cpu().numpy(), G["flow_z"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ld.detach().cpu().numpy(), G["flow_ld"], rtol=2e-4, atol=2e-5)
    nll = (0.5 * torch.sum(z ** 2, 1) - ld).mean()
    nll.backward()
    xo = torch.from_numpy(G["x"]).requires_grad_(True)
    zo, ldo = OF.inn_forward(xo, params)
    OF.nll(zo, ldo).mean().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), xo.grad.numpy(), rtol=2e-3, atol=1e-5)
    with torch.no_grad():
        xr, ldr = inn(z.detach(), rev=True)
    np.testing.assert_allclose(xr.cpu().numpy(), G["x"], rtol=1e-3, atol=2e-5)
    for q in inn.parameters():
        q.requires_grad = True
    # un-frozen flow (the state at every reference call site): autograd now also delivers the parameter gradients,
    # as FrEIA does (train_full_pose_norm_flow.py:75-98 trains the flow through this very call)
    x2 = torch.from_numpy(G["x"]).cuda().requires_grad_(True)
    z2, ld2 = inn(x2)
    (0.5 * torch.sum(z2 ** 2, 1) - ld2).mean().backward()
    np.testing.assert_allclose(x2.grad.cpu().numpy(), xo.grad.numpy(), rtol=2e-3, atol=1e-5)
    from oracle import steps as OS
    pn = OS.params_require_grad(params)
    zo2, ldo2 = OF.inn_forward(torch.from_numpy(G["x"]), pn)
    OF.nll(zo2, ldo2).mean().backward()
    named = dict(inn.named_parameters())
    for k in (0, 3, 7):
        for n, tol in (("subnet.0.weight", 4e-2), ("subnet.0.bias", 4e-2), ("subnet.2.weight", 4e-2), ("subnet.2.bias", 2e-2),
                       ("global_scale", 2e-3), ("global_offset", 2e-3)):
            key = "module_list.%d.%s" % (k, n)
            got, ref = named[key].grad.cpu(), pn[key].grad
            assert got.shape == ref.shape
            e = ((got - ref).norm() / ref.norm()).item()
            assert e < tol, (key, e)
    assert named["module_list.0.w_perm"].grad is None
    # a second call before one backward (data rows + sampled rows in the reference step) accumulates into .grad
    g_before = named["module_list.3.subnet.2.weight"].grad.clone()
    za, lda = inn(x2.detach())
    zb, ldb = inn(x2.detach())
    ((0.5 * torch.sum(za ** 2, 1) - lda).mean() + (0.5 * torch.sum(zb ** 2, 1) - ldb).mean()).backward()
    g_after = named["module_list.3.subnet.2.weight"].grad
    assert ((g_after - 3 * g_before).norm() / g_before.norm()).item() < 1e-3
    # a torch optimiser step changes the parameters: the next call sees the new values (forward AND gradient engine)
    opt = torch.optim.SGD(inn.parameters(), lr=1e-2)
    opt.step()
    opt.zero_grad()
    z3, ld3 = inn(x2.detach())
    (0.5 * torch.sum(z3 ** 2, 1) - ld3).mean().backward()
    p3 = {k: v.detach().cpu().clone() for k, v in inn.state_dict().items()}
    pn3 = OS.params_require_grad(p3)
    zo3, ldo3 = OF.inn_forward(torch.from_numpy(G["x"]), pn3)
    OF.nll(zo3, ldo3).mean().backward()
    np.testing.assert_allclose(z3.detach().cpu().numpy(), zo3.detach().numpy(), rtol=2e-3, atol=2e-4)
    key = "module_list.5.subnet.2.weight"
    e = ((named[key].grad.cpu() - pn3[key].grad).norm() / pn3[key].grad.norm()).item()
    assert e < 4e-2, e
    # opt-out used by code that never reads the flow's gradients (the lifter trainers)
    inn.input_grad_only = True
    opt.zero_grad(set_to_none=True)
    x4 = torch.from_numpy(G["x"]).cuda().requires_grad_(True)
    z4, ld4 = inn(x4)
    (0.5 * torch.sum(z4 ** 2, 1) - ld4).mean().backward()
    assert x4.grad is not None and all(q.grad is None for q in inn.parameters())


def test_metrics_batch_dropin_vs_reference_golden(golden):
    from utils.metrics_batch import Metrics as mb
    G = golden["metrics"]
    gt, pred = torch.from_numpy(G["gt"]).cuda(), torch.from_numpy(G["pred"]).cuda()
    M = gt.shape[0]
    for nj, rj in ((17, 0), (16, 6)):
        g = gt.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj)
        p = pred.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj)
        kw = dict(num_joints=nj, root_joint=rj)
        assert (mb().mpjpe(g, p, **kw).cpu().numpy() - G["mpjpe_j%d_s1" % nj]).__abs__().max() < 0.05
        assert (mb().mpjpe(g, p, use_scaling=False, **kw).cpu().numpy() - G["mpjpe_j%d_s0" % nj]).__abs__().max() < 0.05
        np.testing.assert_allclose(mb().PCK(g, p, **kw).item(), G["pck_j%d" % nj], rtol=1e-6)
        np.testing.assert_allclose(mb().AUC(g, p, **kw).item(), G["auc_j%d" % nj], rtol=1e-5)
        ga = mb().get_all(g, p, **kw)
        for k in ("MPJPE", "PCK", "AUC", "CPS"):
            np.testing.assert_allclose(ga[k].item(), G["getall_%s_j%d" % (k, nj)], rtol=2e-5)
        assert (mb().pmpjpe(g, p, num_joints=nj).cpu().numpy() - G["pmpjpe_batch_j%d" % nj]).__abs__().max() < 0.05
    assert (mb().pmpjpe_best(gt, pred).cpu().numpy() - G["pmpjpe_np_best"]).__abs__().max() < 0.05
    al = mb().procrustes(pred.reshape(M, 3, 17), gt.reshape(M, 3, 17))
    assert al.shape == (M, 3, 17)
    # views that do not start on a 16-byte boundary (row 1 of a [M,51] tensor) are accepted like any other tensor
    assert torch.equal(mb().mpjpe(gt[1:], pred[1:], num_joints=17, root_joint=0), mb().mpjpe(gt, pred, num_joints=17, root_joint=0)[1:])
    assert torch.allclose(mb().pmpjpe_best(gt[1:], pred[1:]), mb().pmpjpe_best(gt, pred)[1:], atol=1e-4)


def test_module_called_twice_before_backward():
    """The reference training_step calls each lifter twice (train_leg_torso_lifter.py:150-151 and :227-228) and each
    occlusion predictor three times (train_occlusion_models.py:196-300) before ONE backward: every call must keep its
    own activations.  Gradients vs the oracle networks on the same inputs."""
    from oracle import nets as ON, steps as OS
    from utils.models_def import Leg_Lifter, Occluded_Limb_Predictor
    g = torch.Generator().manual_seed(5)
    # ---- lifter: two calls, same batch size
    p = ON.init_lifter_params(7, 21)
    m = Leg_Lifter(use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.25).cuda()
    m.load_state_dict(p, strict=False)
    xs = [(torch.randn(96, 14, generator=g) * 0.2) for _ in range(2)]
    xg = [x.cuda().requires_grad_(True) for x in xs]
    outs = [m(x) for x in xg]
    loss = sum((xd.square().sum() + (i + 1) * xa.sum()) for i, (xd, xa) in enumerate(outs))
    loss.backward()
    pr = OS.params_require_grad(p)
    xo = [x.clone().requires_grad_(True) for x in xs]
    ro = [ON.lifter_forward(x, pr) for x in xo]
    sum((xd.square().sum() + (i + 1) * xa.sum()) for i, (xd, xa) in enumerate(ro)).backward()
    for i in range(2):
        assert (outs[i][0].detach().cpu() - ro[i][0].detach()).abs().max() < 5e-3
        assert rel_fro(xg[i].grad.cpu(), xo[i].grad) < 0.1, i
    for name in ("upscale", "res_common.l1", "res_pose2.l2", "res_angle1.l1", "downscale", "angles"):
        mod = m
        for part in name.split("."):
            mod = getattr(mod, part)
        assert rel_fro(mod.weight.grad.cpu(), pr[name + ".weight"].grad) < 8e-2, name
    # the single-call gradient is different (the second call really contributed)
    m.zero_grad()
    xd, xa = m(xg[0].detach())
    (xd.square().sum() + xa.sum()).backward()
    assert rel_fro(m.res_pose2.l2.weight.grad.cpu(), pr["res_pose2.l2.weight"].grad) > 0.2
    # ---- predictor: three calls
    pp = ON.init_predictor_params(14, 9, 33)
    q = Occluded_Limb_Predictor(use_batchnorm=False, num_joints=14).cuda()
    q.load_state_dict(pp, strict=False)
    zs = [torch.randn(64, 42, generator=g) for _ in range(3)]
    ys = [q(z.cuda()) for z in zs]
    sum(((k + 1) * y.square().sum(dim=1).mean()) for k, y in enumerate(ys)).backward()
    pq = OS.params_require_grad(pp)
    sum(((k + 1) * ON.predictor_forward(z, pq).square().sum(dim=1).mean()) for k, z in enumerate(zs)).backward()
    for name in ("upscale", "res_pose1.l1", "res_pose3.l2", "downscale"):
        mod = q
        for part in name.split("."):
            mod = getattr(mod, part)
        assert rel_fro(mod.weight.grad.cpu(), pq[name + ".weight"].grad) < 8e-2, name
    assert q.res_common.l1.weight.grad is None


def test_reference_flow_training_loop_runs_on_the_shim():
    """The reference's own flow trainer body (train_full_pose_norm_flow.py:67-98: `z, jac = inn(x)`, no-grad sampling
    through `inn(z, rev=True)`, `inn(samples)`, `loss.backward()`, `torch.optim.Adam`) executed against the FrEIA shim:
    losses and the parameter trajectory follow the oracle's autograd + Adam on the FrEIA restatement."""
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    from links_b200.synth import synth_poses
    from oracle import flow as OF, steps as OS
    from utils.helpers import subnet_fc
    B = 100                                     # not a multiple of the 64-row GEMM boxes / the 128-row flow tiles
    params = OF.init_flow_params(34, 91, perturb=0.3)
    inn_2d = Ff.SequenceINN(34)
    for _ in range(8):
        inn_2d.append(Fm.AllInOneBlock, subnet_constructor=subnet_fc, permute_soft=True)
    inn_2d.load_state_dict(params)
    inn_2d.cuda()
    optimizer = torch.optim.Adam(inn_2d.parameters(), lr=2e-4, weight_decay=1e-5)
    pn = OS.params_require_grad(params)
    for k in list(pn):
        if "w_perm" in k:
            pn[k].requires_grad_(False)
    opt_ref = torch.optim.Adam([v for v in pn.values() if v.requires_grad], lr=2e-4, weight_decay=1e-5)
    x2d, _ = synth_poses(B, seed=41)
    g = torch.Generator().manual_seed(13)
    for it in range(3):
        x = torch.from_numpy(x2d)
        noise = torch.randn(B, 34, generator=g)
        inp_poses = x.cuda()
        # ---- reference step body
        z, log_jac_det = inn_2d(inp_poses)
        dist_2d = (0.5 * torch.sum(z ** 2, 1) - log_jac_det).mean()
        with torch.no_grad():
            z_noisy = z + 0.2 * noise.cuda() * z                      # add_noise (utils/helpers.py:298-308) with fixed draws
            samples, _ = inn_2d(z_noisy, rev=True)
            samples = samples.reshape(-1, 2, 17)
            samples[:, :, [0]] = 0.0
            samples = samples.reshape(-1, 34)
        z_s, jac_s = inn_2d(samples)
        dist_2d_sample = (0.5 * torch.sum(z_s ** 2, 1) - jac_s).mean()
        loss = dist_2d + dist_2d_sample
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        # ---- oracle
        opt_ref.zero_grad()
        ref = OS.flow_step(x, pn, noise)
        ref["loss"].backward()
        opt_ref.step()
        assert abs(loss.item() - ref["loss"].item()) <= (1e-3 if it == 0 else 5e-3) * abs(ref["loss"].item()), \
            (it, loss.item(), ref["loss"].item())
    named = dict(inn_2d.named_parameters())
    for n in ("subnet.0.weight", "subnet.2.weight", "global_scale", "global_offset"):
        key = "module_list.4." + n
        d_gpu = named[key].detach().cpu() - params[key]
        d_ref = pn[key].detach() - params[key]
        cos = (d_gpu * d_ref).sum() / (d_gpu.norm() * d_ref.norm())
        assert cos.item() > 0.9, (n, cos.item())
    assert torch.equal(named["module_list.4.w_perm"].detach().cpu(), params["module_list.4.w_perm"])


def test_reference_lifter_step_runs_on_the_dropin_modules():
    """The leg/torso training step written the way the reference writes it (train_leg_torso_lifter.py:123-276) -- nn.Module
    lifters called twice, FrEIA flows with requires_grad=True parameters no optimiser owns, utils.helpers /
    rotation_conversions glue, in-place masked assignments, loss.backward(), two torch.optim.Adam -- executed on the
    drop-in modules (CUDA) and compared with the oracle step (CPU fp32): every loss term, the gradient that reaches the
    input of the flows' consumers (through the weights' first Adam update direction)."""
    import math
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    from links_b200.synth import synth_poses
    from oracle import flow as OF, nets as ON, steps as OS
    from utils.helpers import get_bone_lengths_all, perspective_projection, subnet_fc
    from utils.models_def import Leg_Lifter, Torso_Lifter
    from utils.rotation_conversions import euler_angles_to_matrix
    dev = "cuda"
    B, depth_t = 48, 10.0
    p_leg, p_torso = ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)
    f_leg, f_torso = OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3)
    f_full = OF.init_flow_params(34, 40, perturb=0.3)
    legs_lifter = Leg_Lifter(use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.25).cuda()
    torso_lifter = Torso_Lifter(use_batchnorm=False, num_joints=10, use_dropout=False, d_rate=0.25).cuda()
    legs_lifter.load_state_dict(p_leg, strict=False)
    torso_lifter.load_state_dict(p_torso, strict=False)

    def make_inn(C, params):
        inn = Ff.SequenceINN(C)
        for _ in range(8):
            inn.append(Fm.AllInOneBlock, subnet_constructor=subnet_fc, permute_soft=True)
        inn.load_state_dict(params)
        return inn.cuda()
    leg_inn_2d, torso_inn_2d, inn_2d = make_inn(14, f_leg), make_inn(20, f_torso), make_inn(34, f_full)
    leg_opt = torch.optim.Adam(legs_lifter.parameters(), lr=2e-4, weight_decay=1e-5)
    torso_opt = torch.optim.Adam(torso_lifter.parameters(), lr=2e-4, weight_decay=1e-5)
    x2d, _ = synth_poses(B, seed=3)
    g = torch.Generator().manual_seed(8)
    x, noise = torch.from_numpy(x2d), torch.randn(B, 34, generator=g)
    eps_x, u_y = torch.randn(2 * B, generator=g), torch.rand(2 * B, generator=g)
    cfg = OS.DEFAULT_CFG
    bone_rel = torch.tensor(OS.G.BONE_REL_MPI, dtype=torch.float32)

    # ---------------- the step, reference style, on the drop-in modules
    leg_opt.zero_grad()
    torso_opt.zero_grad()
    inp_poses = x.to(dev)
    with torch.no_grad():                                               # :133-142
        z, _ = inn_2d(inp_poses)
        z = z + 0.2 * noise.to(dev) * z
        sampled_poses, _ = inn_2d(z, rev=True)
        sampled_poses = sampled_poses.reshape(-1, 2, 17)
        sampled_poses[:, :, [0]] = 0.0
        sampled_poses = sampled_poses.reshape(-1, 34)
    inp_poses = torch.cat((inp_poses, sampled_poses.data), dim=0)
    n = inp_poses.shape[0]
    inp_legs = inp_poses.reshape(-1, 2, 17)[:, :, :7].reshape(-1, 14)
    inp_torso = inp_poses.reshape(-1, 2, 17)[:, :, 7:].reshape(-1, 20)
    legs_pred, legs_angle = legs_lifter(inp_legs)
    torso_pred, torso_angle = torso_lifter(inp_torso)
    props = (legs_angle + torso_angle) / 2
    pred = torch.cat((legs_pred, torso_pred), dim=1)
    pred[:, 0] = 0.0
    zeros = torch.zeros((n, 1), device=dev)
    R_comp = euler_angles_to_matrix(torch.cat((torch.ones((n, 1), device=dev) * props, zeros, zeros), dim=1), 'XYZ')
    elevation = torch.cat((props.mean().reshape(1), props.std().reshape(1)))
    x_ang = (-elevation[0]) + elevation[1] * eps_x.to(dev).reshape(n, 1)
    y_ang = (u_y.to(dev).reshape(n, 1) - 0.5) * 1.99 * math.pi
    Rx = euler_angles_to_matrix(torch.cat((x_ang, zeros, zeros), dim=1), 'XYZ')
    Ry = euler_angles_to_matrix(torch.cat((zeros, y_ang, zeros), dim=1), 'XYZ')
    R = Rx @ (Ry @ R_comp)
    depth = pred + depth_t
    depth[depth < 1.0] = 1.0
    pred_3d = torch.cat(((inp_poses.reshape(-1, 2, 17) * depth.reshape(-1, 1, 17).repeat(1, 2, 1)).reshape(-1, 34), depth),
                        dim=1).reshape(-1, 3, 17)
    pred_3d = pred_3d - pred_3d[:, :, [0]]
    rot_poses = (R.matmul(pred_3d)).reshape(-1, 51)
    rot_2d = perspective_projection(torch.cat((rot_poses[:, 0:34], rot_poses[:, 34:51] + depth_t), dim=1))
    torso_norm = rot_2d.reshape(-1, 2, 17)[:, :, 7:].reshape(-1, 20)
    leg_norm = rot_2d.reshape(-1, 2, 17)[:, :, :7].reshape(-1, 14)
    z, jac = leg_inn_2d(leg_norm)
    leg_likeli = (0.5 * torch.sum(z ** 2, 1) - jac).mean()
    z, jac = torso_inn_2d(torso_norm)
    torso_likeli = (0.5 * torch.sum(z ** 2, 1) - jac).mean()
    likeli = torso_likeli + leg_likeli
    legs_pred_rot, _ = legs_lifter(leg_norm)
    torso_pred_rot, _ = torso_lifter(torso_norm)
    pred_rot = torch.cat((legs_pred_rot, torso_pred_rot), dim=1)
    pred_rot[:, 0] = 0.0
    pred_rot_depth = pred_rot + depth_t
    pred_rot_depth[pred_rot_depth < 1.0] = 1.0
    pred_3d_rot = torch.cat(((rot_2d.reshape(-1, 2, 17) * pred_rot_depth.reshape(-1, 1, 17).repeat(1, 2, 1)).reshape(-1, 34),
                             pred_rot_depth), dim=1).reshape(-1, 3, 17)
    pred_3d_rot = pred_3d_rot - pred_3d_rot[:, :, [0]]
    L3d = (rot_poses - pred_3d_rot.reshape(-1, 51)).norm(dim=1).mean()
    re_rot_3d = (R.permute(0, 2, 1) @ pred_3d_rot).reshape(-1, 51)
    re_rot_2d = perspective_projection(torch.cat((re_rot_3d[:, 0:34], re_rot_3d[:, 34:51] + depth_t), dim=1))
    rep_rot = (re_rot_2d - inp_poses).abs().sum(dim=1).mean()
    num_pairs = n // 2
    pose_pairs = pred_3d[0:2 * num_pairs].reshape(2 * num_pairs, 51).reshape(-1, 2, 51)
    pairs_re = re_rot_3d[0:2 * num_pairs].reshape(-1, 2, 51)
    vel = ((pose_pairs[:, 0] - pose_pairs[:, 1]) - (pairs_re[:, 0] - pairs_re[:, 1])).norm(dim=1).mean()
    bl = get_bone_lengths_all(pred_3d.reshape(-1, 51))
    rel_bl = bl / bl.mean(dim=1, keepdim=True)
    bl_prior = (bone_rel.to(dev) - rel_bl).square().sum(dim=1).mean()
    loss = cfg["weight_likeli"] * likeli + cfg["weight_2d"] * rep_rot + cfg["weight_3d"] * L3d + \
        cfg["weight_velocity"] * vel + cfg["weight_bl"] * bl_prior
    loss.backward()
    torso_opt.step()
    leg_opt.step()
    got = {"L3d": L3d, "rep_rot": rep_rot, "re_rot_3d": vel, "bl_prior": bl_prior, "leg_likeli": leg_likeli,
           "torso_likeli": torso_likeli, "likeli": likeli, "loss": loss}

    # ---------------- the oracle step (CPU fp32) on the same inputs
    pl, pt = OS.params_require_grad(p_leg), OS.params_require_grad(p_torso)
    opts = OS.make_adam([pl, pt])
    u = OS.sample_poses(x, f_full, noise)
    ref = OS.lt_step(u, pl, pt, f_leg, f_torso, eps_x, u_y, cfg, bone_rel)
    ref["loss"].backward()
    for o in opts:
        o.step()
    for k, v in got.items():
        r = ref[k].item()
        assert abs(v.item() - r) <= 2e-3 * abs(r) + 1e-6, (k, v.item(), r)
    # first Adam update = -lr * sign(gradient) (up to eps): the two implementations moved the weights the same way
    for mod, pd, p0 in ((legs_lifter, pl, p_leg), (torso_lifter, pt, p_torso)):
        named = dict(mod.named_parameters())
        for key in ("upscale.weight", "res_pose2.l1.weight", "res_angle1.l2.weight", "downscale.weight", "angles.bias"):
            d_gpu = (named[key].detach().cpu() - p0[key]).flatten()
            d_ref = (pd[key].detach() - p0[key]).flatten()
            agree = (torch.sign(d_gpu) == torch.sign(d_ref)).float().mean().item()
            assert agree > 0.9, (key, agree)
    # the flows' parameters got gradients too (requires_grad=True, like in the reference), nobody steps them
    assert next(leg_inn_2d.module_list[0].subnet.parameters()).grad is not None
