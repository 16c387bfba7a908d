"""GPU parity AT THE BASELINE.json CONFIG SIZES (VERDICT r1, weak #1b): the multi-tile GEMM scheduling, the two-pass
wgrad contraction (K = 2N rows) and the 148-CTA paths are exercised through whole steps against the CPU oracle, not
only through GEMM unit tests.

  config #2  LT and LR lifter step, B = 1024 poses (N = 2048 rows)       train_leg_torso_lifter.py:123-284,
  config #3  same step at B = 8192 (N = 16384 rows)                       train_left_right_lifter.py:121-435
  config #4  occlusion step, B = 512 poses per GPU (4096 / 8 GPUs)        train_occlusion_models.py:144-314
  config #5  eval: lift + N-MPJPE + PA-MPJPE over 204 800 poses           eval_h36m.py:50-97

Tolerances (north star): losses and joints 1e-3 relative (bf16 operands, fp32 accumulate), MPJPE / PA-MPJPE 0.05 mm.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def _weights(kind):
    from oracle import flow as OF, nets as ON
    if kind == "lt":
        nets = [ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)]
        flows = [OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3)]
    else:
        nets = [ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)]
        flows = [OF.init_flow_params(22, 43, perturb=0.3), OF.init_flow_params(22, 44, perturb=0.3)]
    return nets, flows, OF.init_flow_params(34, 40, perturb=0.3)


@pytest.mark.parametrize("kind,B", [("lt", 1024), ("lr", 1024), ("lt", 8192), ("lr", 8192)])
def test_lifter_step_at_baseline_batch(kind, B):
    from links_b200.steps import LifterStep
    from links_b200.synth import synth_poses
    from oracle import steps as OS
    nets, flows, full = _weights(kind)
    step = LifterStep(kind, B, nets, flows, full)
    x2d, _ = synth_poses(B, seed=500 + B)
    g = torch.Generator().manual_seed(B)
    x, noise = torch.from_numpy(x2d), torch.randn(B, 34, generator=g)
    eps_x, u_y = torch.randn(2 * B, generator=g), torch.rand(2 * B, generator=g)
    step.x.copy_(x); step.noise.copy_(noise); step.eps_x.copy_(eps_x); step.u_y.copy_(u_y)
    step.forward_backward()
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    pn = [OS.params_require_grad(p) for p in nets]
    u = OS.sample_poses(x, full, noise)
    aux = {}
    fn = OS.lt_step if kind == "lt" else OS.lr_step
    ref = fn(u, pn[0], pn[1], flows[0], flows[1], eps_x, u_y, aux=aux)
    ref["loss"].backward()
    # sampled poses, projected joints, losses: 1e-3 relative
    assert rel_fro(step.u.cpu(), u) < 1e-3
    key = "rot_2d" if kind == "lt" else "rot_2d_left"
    assert rel_fro(step.qfull[0].cpu(), aux[key].detach()) < 1e-3
    got = step.loss_dict()
    for k, v in got.items():
        r = ref[k].item()
        assert abs(v - r) <= 1e-3 * abs(r) + 1e-6, (kind, B, k, v, r)
    # weight gradients (two-pass wgrad contraction over 2N rows) at bf16-operand accuracy
    for s in range(2):
        for name in ("upscale", "res_common.l2", "res_pose3.l1", "res_angle1.l2", "downscale", "angles"):
            e = rel_fro(step.mlp.nets[s].layers[name].gW.cpu(), pn[s][name + ".weight"].grad)
            assert e < 6e-2, (kind, B, s, name, e)
            eb = rel_fro(step.mlp.nets[s].layers[name].gb.cpu(), pn[s][name + ".bias"].grad)
            assert eb < 6e-2, (kind, B, s, name, eb)


def test_occlusion_step_at_baseline_batch():
    """config #4: 4096 poses over 8 GPUs = 512 per GPU, 3 rounds -> 1536 GEMM rows per predictor."""
    from links_b200.occlusion import OCC_IN, OCC_NAMES, OCC_OUT, OcclusionStep
    from links_b200.synth import synth_poses
    from oracle import nets as ON, steps as OS
    B = 512
    lifters = [ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)]
    preds = {n: ON.init_predictor_params(OCC_IN[n] // 3, OCC_OUT[n], 100 + i) for i, n in enumerate(OCC_NAMES)}
    step = OcclusionStep(B, lifters, preds)
    x2d, _ = synth_poses(B, seed=19)
    g = torch.Generator().manual_seed(13)
    u1, u2 = torch.rand(B, generator=g), torch.rand(B, generator=g)
    x = torch.from_numpy(x2d)
    step.x.copy_(x); step.u_y[0].copy_(u1); step.u_y[1].copy_(u2)
    step.forward_backward()
    torch.cuda.synchronize()
    pn = {n: OS.params_require_grad(p) for n, p in preds.items()}
    ref = OS.occlusion_step(x, lifters[0], lifters[1], pn, u1, u2)
    ref["loss"].backward()
    got = step.loss_dict()
    for k, v in got.items():
        r = ref[k].item()
        assert abs(v - r) <= 1e-3 * abs(r), (k, v, r)
    for s, n in enumerate(OCC_NAMES):
        for name in ("upscale", "res_pose1.l2", "res_pose3.l1", "downscale"):
            e = rel_fro(step.mlp.nets[s].layers[name].gW.cpu(), pn[n][name + ".weight"].grad)
            assert e < 8e-2, (n, name, e)


def test_eval_at_200k_poses():
    """config #5 (per-GPU slice, bounded so the CPU oracle finishes in about a minute): 204 800 poses through the
    sharded eval path in 65 536-pose chunks (full chunks + a ragged tail) vs the oracle's eval_h36m restatement."""
    from links_b200.occlusion import EvalRunner
    from links_b200.synth import synth_poses
    from oracle import nets as ON, steps as OS
    n, chunk = 204_800, 65_536
    p2d, gt = synth_poses(n, seed=77)
    params = [ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)]
    ev = EvalRunner("lr", params, chunk=chunk)
    xd, gd = torch.from_numpy(p2d).cuda(), torch.from_numpy(gt).cuda()
    for i in range(0, n, chunk):
        ev.run_chunk(xd[i:i + chunk].contiguous(), gd[i:i + chunk].contiguous())
    out = ev.result()
    assert out["count"] == n
    pa, nm, cnt = 0.0, 0.0, 0
    for i in range(0, n, 16384):                       # oracle in slices (keeps the CPU working set small)
        xs, gs = torch.from_numpy(p2d[i:i + 16384]), torch.from_numpy(gt[i:i + 16384])
        pred = OS.eval_lr_predict(xs, params[0], params[1], choice="right")
        m = OS.eval_metrics(gs, pred)
        pa += m["pa_mpjpe"] * xs.shape[0]; nm += m["n_mpjpe"] * xs.shape[0]; cnt += xs.shape[0]
    assert cnt == n
    assert abs(out["pa_mpjpe"] - pa / n) < 0.05, (out, pa / n)
    assert abs(out["n_mpjpe"] - nm / n) < 0.05, (out, nm / n)


def test_merged_lt_lr_step_with_sampling_prefetch():
    """kind='both' (config #2 as bench.py runs it): the leg/torso and the left/right step of one batch share a
    4-network engine, one sampling pass and -- with prefetch_sample -- draw the poses of step k+1 while step k runs.
    Both flavours' losses and gradients vs the oracle, for two consecutive batches (the second one exercises the
    prefetched poses)."""
    from links_b200.steps import LifterStep
    from links_b200.synth import synth_poses
    from oracle import steps as OS
    B = 1024
    n_lt, f_lt, full = _weights("lt")
    n_lr, f_lr, _ = _weights("lr")
    step = LifterStep("both", B, n_lt + n_lr, f_lt + f_lr, full, cfg={"prefetch_sample": True})
    batches = []
    for i in range(2):
        x2d, _ = synth_poses(B, seed=900 + i)
        g = torch.Generator().manual_seed(40 + i)
        batches.append(dict(x=torch.from_numpy(x2d), noise=torch.randn(B, 34, generator=g),
                            eps_x=torch.randn(2 * B, generator=g), u_y=torch.rand(2 * B, generator=g)))
    step.x.copy_(batches[0]["x"]); step.noise.copy_(batches[0]["noise"])
    step.prime()
    for i in range(2):
        nxt = batches[min(i + 1, 1)]
        step.x.copy_(nxt["x"]); step.noise.copy_(nxt["noise"])                     # sampling inputs of the NEXT step
        step.eps_x.copy_(batches[i]["eps_x"]); step.u_y.copy_(batches[i]["u_y"])    # rotation draws of THIS step
        step.forward_backward()
        torch.cuda.synchronize()
        u = OS.sample_poses(batches[i]["x"], full, batches[i]["noise"])
        assert rel_fro(step.u.cpu(), u) < 1e-3
        got = step.loss_dict()
        for kind, nets, flows, fn, s0 in (("lt", n_lt, f_lt, OS.lt_step, 0), ("lr", n_lr, f_lr, OS.lr_step, 2)):
            pn = [OS.params_require_grad(p) for p in nets]
            ref = fn(u, pn[0], pn[1], flows[0], flows[1], batches[i]["eps_x"], batches[i]["u_y"])
            for k, v in got[kind].items():
                r = ref[k].item()
                assert abs(v - r) <= 1e-3 * abs(r) + 1e-6, (i, kind, k, v, r)
            if i == 1:
                ref["loss"].backward()
                for s in range(2):
                    for name in ("upscale", "res_common.l1", "res_pose2.l2", "res_angle3.l1", "downscale", "angles"):
                        e = rel_fro(step.mlp.nets[s0 + s].layers[name].gW.cpu(), pn[s][name + ".weight"].grad)
                        assert e < 6e-2, (kind, s, name, e)
