"""GPU: flow training step (train_full_pose_norm_flow.py:67-98) -- loss, every parameter gradient and the trajectory
after a few Adam steps vs the oracle's autograd on the FrEIA restatement."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_flow_train_step_vs_oracle():
    from links_b200.flowtrain import FlowTrainStep
    from links_b200.synth import synth_poses
    from oracle import flow as OF, steps as OS
    B, C = 96, 34
    params = OF.init_flow_params(C, 77, perturb=0.3)
    step = FlowTrainStep(C, params, B)
    x2d, _ = synth_poses(B, seed=17)
    g = torch.Generator().manual_seed(5)
    x, noise = torch.from_numpy(x2d), torch.randn(B, C, generator=g)
    pn = OS.params_require_grad(params)
    for k in list(pn):
        if "w_perm" in k:
            pn[k].requires_grad_(False)
    opt = torch.optim.Adam([v for v in pn.values() if v.requires_grad], lr=2e-4, weight_decay=1e-5)
    step.wd = 1e-5
    for it in range(3):
        step.x.copy_(x); step.noise.copy_(noise)
        step.forward_backward()
        torch.cuda.synchronize()
        opt.zero_grad()
        out = OS.flow_step(x, pn, noise)
        out["loss"].backward()
        ref_loss = out["loss"].item()
        got = step.loss_dict()["loss"]
        assert abs(got - ref_loss) <= (1e-3 if it == 0 else 5e-3) * abs(ref_loss), (it, got, ref_loss)
        if it == 0:
            for k in range(8):
                for n, tol in (("subnet.0.weight", 4e-2), ("subnet.0.bias", 4e-2), ("subnet.2.weight", 4e-2),
                               ("subnet.2.bias", 2e-2), ("global_scale", 1e-3), ("global_offset", 1e-3)):
                    ref = pn["module_list.%d.%s" % (k, n)].grad
                    e = rel_fro(step.Gd[k][n].cpu(), ref)
                    assert e < tol, (k, n, e)
        step.optimizer_step()
        opt.step()
    # the two trajectories moved the parameters the same way
    for n in ("subnet.0.weight", "subnet.2.weight", "global_offset"):
        key = "module_list.3." + n
        d_gpu = step.P[3][n].cpu() - params[key]
        d_ref = pn[key].detach() - params[key]
        cos = (d_gpu * d_ref).sum() / (d_gpu.norm() * d_ref.norm())
        assert cos.item() > 0.9, (n, cos.item())


def test_part_flow_trainer_vs_oracle():
    """train_leg_torso_left_right_norm_flow.py:100-174: four part flows on data + samples of the frozen full flow."""
    from links_b200.flowtrain import PartFlowTrainer
    from links_b200.synth import synth_poses
    from oracle import flow as OF, steps as OS
    B = 64
    full = OF.init_flow_params(34, 40, perturb=0.3)
    width = {"legs": 14, "torso": 20, "left": 22, "right": 22}
    parts = {n: OF.init_flow_params(width[n], 60 + i, perturb=0.3) for i, n in enumerate(PartFlowTrainer.NAMES)}
    tr = PartFlowTrainer(full, parts, B)
    x2d, _ = synth_poses(B, seed=23)
    g = torch.Generator().manual_seed(9)
    x, noise = torch.from_numpy(x2d), torch.randn(B, 34, generator=g)
    pn = {n: OS.params_require_grad(parts[n]) for n in parts}
    ref = OS.part_flow_step(x, full, pn, noise)
    ref["loss"].backward()
    tr.x.copy_(x); tr.noise.copy_(noise)
    for st in tr.steps.values():        # forward/backward only, keep the parameters for the gradient comparison
        st.optimizer_step = lambda: None
    tr.step()
    torch.cuda.synchronize()
    got = tr.loss_dict()
    for n in PartFlowTrainer.NAMES:
        r = (ref["dist_2d_" + n] + ref["dist_2d_%s_sample" % n]).item()
        assert abs(got["dist_2d_" + n] - r) <= 1e-3 * abs(r), (n, got["dist_2d_" + n], r)
        e = rel_fro(tr.steps[n].Gd[2]["subnet.2.weight"].cpu(), pn[n]["module_list.2.subnet.2.weight"].grad)
        assert e < 4e-2, (n, e)
    assert abs(got["loss"] - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())


@pytest.mark.parametrize("which", ["full", "parts"])
def test_graph_replay_equals_eager_steps(which):
    """run() (eager first step, then one captured CUDA graph replayed; what the drop-in scripts call) follows the same
    trajectory as step() launched eagerly: same losses every step, same parameters after five steps.  Also pins the
    all-blocks strided views (shadow casts, global-affine gradient slots) against the per-block parameter views."""
    from links_b200.flowtrain import FlowTrainStep, PartFlowTrainer
    from links_b200.synth import synth_poses
    from oracle import flow as OF
    B = 64
    x2d, _ = synth_poses(B, seed=31)
    g = torch.Generator().manual_seed(11)
    xs = torch.from_numpy(x2d).cuda()
    noises = [torch.randn(B, 34, generator=g).cuda() for _ in range(5)]

    def make():
        if which == "full":
            return FlowTrainStep(34, OF.init_flow_params(34, 77, perturb=0.3), B, weight_decay=1e-5)
        width = {"legs": 14, "torso": 20, "left": 22, "right": 22}
        parts = {n: OF.init_flow_params(width[n], 60 + i, perturb=0.3) for i, n in enumerate(PartFlowTrainer.NAMES)}
        return PartFlowTrainer(OF.init_flow_params(34, 40, perturb=0.3), parts, B)

    def masters(t):
        return [t.master] if which == "full" else [t.steps[n].master for n in PartFlowTrainer.NAMES]

    a, b = make(), make()
    for it in range(5):
        for t in (a, b):
            t.x.copy_(xs); t.noise.copy_(noises[it])
        a.step()
        b.run()
        la, lb = a.loss_dict()["loss"], b.loss_dict()["loss"]
        assert np.isfinite(la) and abs(la - lb) <= 2e-5 * abs(la), (it, la, lb)       # atomics: summation order only
    assert b.graph is not None
    for ma, mb in zip(masters(a), masters(b)):
        d = (ma - mb).abs()
        # atomically accumulated global-affine gradients differ in their last bits between runs; Adam turns that into at
        # most a fraction of one update (lr = 2e-4) on entries whose gradient is ~0
        assert d.max().item() <= 1e-4 and d.mean().item() <= 1e-6, (d.max().item(), d.mean().item())
    if which == "full":
        # the strided all-blocks views are the per-block views
        for k in (0, 5):
            assert a._w1_all[k].data_ptr() == a.P[k]["subnet.0.weight"].data_ptr()
            assert a._dgo_all[k].data_ptr() == a.Gd[k]["global_offset"].data_ptr()
            assert torch.equal(a.W1b[k, :, :a.c1].float(), a.P[k]["subnet.0.weight"].bfloat16().float())
            assert torch.equal(a.W2b[k, :2 * a.c2].float(), a.P[k]["subnet.2.weight"].bfloat16().float())
