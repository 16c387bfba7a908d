"""Worker of tests/test_gpu_dp.py: run under `python -m torch.distributed.run --nproc-per-node W` on W GPUs.

Checks the PRODUCT data-parallel path (LifterStep._on_bucket: [bf16 compress ->] NCCL all-reduce -> Adam, overlapped
with backward) against
  (1) the same engine without communication: reduced gradients == sum over ranks of the local gradients
      (fp32 buckets: to fp32 summation-order noise; bf16 buckets: to bf16 rounding of each addend and partial sum);
  (2) the CPU oracle run on every rank's shard (per-shard elevation statistics = what DDP of the reference computes,
      train_leg_torso_lifter.py:168): averaged oracle gradients vs the reduced gradients, and the direction of the
      first Adam update.
Usage: dp_worker.py <kind> <grad_comm> <B_per_rank> <out_dir>"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def main():
    kind, grad_comm, B, out_dir = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from links_b200.shard import shard_rows
    from links_b200.steps import LifterStep
    from links_b200.synth import synth_poses
    from oracle import flow as OF, nets as ON, steps as OS
    if kind == "lt":
        nets = [ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)]
        flows = [OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3)]
    else:
        nets = [ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)]
        flows = [OF.init_flow_params(22, 43, perturb=0.3), OF.init_flow_params(22, 44, perturb=0.3)]
    full = OF.init_flow_params(34, 40, perturb=0.3)
    # the GLOBAL batch and draws are generated identically on every rank, then sharded (even contiguous slices)
    Bg = B * world
    x2d, _ = synth_poses(Bg, seed=91)
    g = torch.Generator().manual_seed(17)
    X, NOISE = torch.from_numpy(x2d), torch.randn(Bg, 34, generator=g)
    # per-row draws for the [real ; sampled] halves of every shard
    EPS = torch.randn(world, 2 * B, generator=g)
    UY = torch.rand(world, 2 * B, generator=g)

    def load(step, r):
        step.x.copy_(shard_rows(X, r, world)); step.noise.copy_(shard_rows(NOISE, r, world))
        step.eps_x.copy_(EPS[r]); step.u_y.copy_(UY[r])

    # ---- (1) local gradients of this rank's shard, no communication
    solo = LifterStep(kind, B, nets, flows, full, cfg={"dp_buckets": 1 if grad_comm.startswith("push") else 2, "dp_layout": True})   # same flat layout
    load(solo, rank)
    solo.forward_backward()
    torch.cuda.synchronize()
    g_local = solo.mlp.grad.clone()
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)                                   # fp32 reference sum over ranks
    abs_sum = g_local.abs()
    dist.all_reduce(abs_sum)

    # ---- the product path
    gstats = grad_comm == "push_global"
    if gstats:
        grad_comm = "push"
    cfg = {"grad_comm": grad_comm, "dp_buckets": 2, "global_elevation_stats": gstats}
    step = LifterStep(kind, B, nets, flows, full, cfg=cfg, process_group=dist.group.WORLD)
    load(step, rank)
    w0 = step.mlp.master.clone()
    step.step()
    torch.cuda.synchronize()
    m = step.mlp
    if gstats:
        return check_global(step, nets, flows, full, X, NOISE, EPS, UY, rank, world, kind, out_dir)
    if grad_comm == "push":
        return check_push(step, solo, nets, flows, full, X, NOISE, EPS, UY, g_sum, abs_sum, w0, rank, world, kind, out_dir)
    reduced = m.grad16.float() if grad_comm == "bf16" else m.grad
    err = (reduced - g_sum).abs()
    if grad_comm == "bf16":
        # each addend is rounded to bf16 (2^-9 relative) and so is every partial sum of the ring / tree
        bound = abs_sum * (2.0 ** -8) * world + 1e-30     # bf16 unit roundoff 2^-8: every addend once, every partial sum once
        worst = (err / bound).max().item()
        assert worst <= 1.0, "bf16 bucket all-reduce outside bf16 rounding: %g" % worst
        assert rel_fro(reduced, g_sum) < 4e-3
    else:
        assert rel_fro(reduced, g_sum) < 1e-6
    # every rank holds identical parameters after the step
    chk = torch.stack((m.master.double().sum(), m.master.double().abs().sum()))
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks diverged after one data-parallel step"

    # ---- (2) the oracle on every shard (rank 0 only: CPU work)
    if rank == 0:
        pn = [OS.params_require_grad(p) for p in nets]
        fn = OS.lt_step if kind == "lt" else OS.lr_step
        loss_mean = {}
        for r in range(world):
            u = OS.sample_poses(shard_rows(X, r, world), full, shard_rows(NOISE, r, world))
            out = fn(u, pn[0], pn[1], flows[0], flows[1], EPS[r], UY[r])
            (out["loss"] / world).backward()                 # DDP: mean of the per-shard mean losses
            for k, v in out.items():
                loss_mean[k] = loss_mean.get(k, 0.0) + v.item() / world
        for s in range(2):
            for name in ("upscale", "res_common.l1", "res_pose2.l2", "res_angle3.l1", "downscale", "angles"):
                L = m.nets[s].layers[name]
                off = (L.gW.data_ptr() - m.grad.data_ptr()) // 4
                red = reduced[off:off + L.N * L.K].view(L.N, L.K).cpu() / world
                e = rel_fro(red, pn[s][name + ".weight"].grad)
                assert e < 6e-2, (s, name, e)
        opts = OS.make_adam(pn)
        for o in opts:
            o.step()
        for s in range(2):
            L = m.nets[s].layers["res_pose1.l1"]
            d_gpu = L.W.cpu() - nets[s]["res_pose1.l1.weight"]
            d_ref = pn[s]["res_pose1.l1.weight"].detach() - nets[s]["res_pose1.l1.weight"]
            cos = ((d_gpu * d_ref).sum() / (d_gpu.norm() * d_ref.norm())).item()
            assert cos > 0.9, cos
        # this rank's losses are its shard's losses
        u = OS.sample_poses(shard_rows(X, 0, world), full, shard_rows(NOISE, 0, world))
        pn0 = [OS.params_require_grad(p, False) for p in nets]
        with torch.no_grad():
            out0 = fn(u, pn0[0], pn0[1], flows[0], flows[1], EPS[0], UY[0])
        for k, v in step.loss_dict().items():
            assert abs(v - out0[k].item()) <= 1e-3 * abs(out0[k].item()) + 1e-6, (k, v, out0[k].item())
    assert (m.master - w0).abs().max().item() > 0
    dist.barrier()
    open(os.path.join(out_dir, "ok_%s_%s_%d" % (kind, grad_comm, rank)), "w").write("ok")
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)      # captured/in-flight NCCL state: leave without running destructors (see bench.py)


def check_push(step, solo, nets, flows, full, X, NOISE, EPS, UY, g_sum, abs_sum, w0, rank, world, kind, out_dir):
    """grad_comm = "push": reduce-scatter by peer stores out of the wgrad epilogues, sharded Adam, shadows stored into every
    rank.  (1) the W staging slots of this rank hold every rank's bf16 gradient of the rows it owns: their fp32 sum equals
    the sum of the ranks' local gradients to bf16 rounding of each addend; (2) every rank ends the step with identical
    bf16 shadows, equal to bf16(master) on the owner; (3) the owned master rows moved the way the oracle's Adam moves them."""
    from links_b200.shard import shard_rows
    from oracle import steps as OS
    m, ms = step.mlp, solo.mlp
    z = m.zero
    rpo, cols, owned, W = z["rpo"], z["cols"], z["owned"], z["W"]
    stage = z["stage"].view(W, owned).float()
    worst = 0.0
    for li, (s, n) in enumerate(z["big"]):
        Ls = ms.nets[s].layers[n]                       # same layer in the un-communicated engine (its own flat layout)
        blk = slice(li * rpo * cols, (li + 1) * rpo * cols)
        got = stage[:, blk].sum(0).view(rpo, cols)
        off = Ls.off_W + rank * rpo * cols
        ref = g_sum[off:off + rpo * cols].view(rpo, cols)
        bound = abs_sum[off:off + rpo * cols].view(rpo, cols) * 2.0 ** -8 + 1e-30
        worst = max(worst, ((got - ref).abs() / bound).max().item())
    assert worst <= 1.0, "push reduce-scatter outside bf16 rounding: %g" % worst
    # rest (small layers, biases): plain NCCL sum in fp32
    a, e = m.bucket_mid[0], m.bucket_ranges[0][1]
    a_s, e_s = ms.bucket_mid[0], ms.bucket_ranges[0][1]
    assert (e - a) == (e_s - a_s)
    assert rel_fro(m.grad[a:e], g_sum[a_s:e_s]) < 1e-6
    # identical shadows everywhere, = bf16(master) on the owner's rows
    chk = torch.stack((z["shadow"].double().sum(), z["shadow"].double().abs().sum()))
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks hold different shadows after a push step"
    for s, n in z["big"][:6]:
        L = m.nets[s].layers[n]
        own = slice(rank * rpo, (rank + 1) * rpo)
        assert torch.equal(L.Wb[own], L.W[own].bfloat16()), (s, n)
    assert (m.master - w0).abs().max().item() > 0
    # complete masters after the explicit gather; then the direction of the first Adam update vs the oracle
    m.zero_sync_master()
    chk = torch.stack((m.master.double().sum(), m.master.double().abs().sum()))
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "masters differ after zero_sync_master"
    if rank == 0:
        pn = [OS.params_require_grad(p) for p in nets]
        fn = OS.lt_step if kind == "lt" else OS.lr_step
        for r in range(world):
            u = OS.sample_poses(shard_rows(X, r, world), full, shard_rows(NOISE, r, world))
            out = fn(u, pn[0], pn[1], flows[0], flows[1], EPS[r], UY[r])
            (out["loss"] / world).backward()
        opts = OS.make_adam(pn)
        for o in opts:
            o.step()
        for s in range(2):
            for name in ("res_pose1.l1", "res_angle2.l2", "downscale", "upscale"):
                L = m.nets[s].layers[name]
                d_gpu = L.W.cpu() - nets[s][name + ".weight"]
                d_ref = pn[s][name + ".weight"].detach() - nets[s][name + ".weight"]
                cos = ((d_gpu * d_ref).sum() / (d_gpu.norm() * d_ref.norm())).item()
                assert cos > 0.9, (s, name, cos)
    dist.barrier()
    open(os.path.join(out_dir, "ok_%s_%s_%d" % (kind, "push", rank)), "w").write("ok")
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


def check_global(step, nets, flows, full, X, NOISE, EPS, UY, rank, world, kind, out_dir):
    """global_elevation_stats: the data-parallel step must equal the reference step on the CONCATENATED batch (the only
    term of the step that couples rows across ranks is props.mean() / props.std(), train_leg_torso_lifter.py:168):
    rank-averaged losses, the reduced gradients of the owned rows and the first Adam update vs ONE oracle step on all rows."""
    from links_b200.shard import shard_rows
    from oracle import steps as OS
    m = step.mlp
    z = m.zero
    losses = torch.tensor(list(step.loss_dict().values()), dtype=torch.float64, device="cuda")
    dist.all_reduce(losses)
    losses = (losses / world).tolist()
    m.zero_sync_master()
    if rank == 0:
        pn = [OS.params_require_grad(p) for p in nets]
        fn = OS.lt_step if kind == "lt" else OS.lr_step
        us = [OS.sample_poses(shard_rows(X, r, world), full, shard_rows(NOISE, r, world)) for r in range(world)]
        u = torch.cat(us, dim=0)                     # every shard = [real ; sampled]: row pairs stay inside a shard
        out = fn(u, pn[0], pn[1], flows[0], flows[1], EPS.reshape(-1), UY.reshape(-1))
        out["loss"].backward()
        for (k, v), got in zip(step.loss_dict().items(), losses):
            r_ = out[k].item()
            assert abs(got - r_) <= 1e-3 * abs(r_) + 1e-6, (k, got, r_)
        rpo, cols, owned, W = z["rpo"], z["cols"], z["owned"], z["W"]
        stage = z["stage"].view(W, owned).float()
        for li, (s, n) in enumerate(z["big"]):
            if n not in ("res_common.l1", "res_pose2.l2", "res_angle1.l1", "res_angle3.l2"):
                continue
            got = stage[:, li * rpo * cols:(li + 1) * rpo * cols].sum(0).view(rpo, cols).cpu() / world
            ref = pn[s][n + ".weight"].grad[rank * rpo:(rank + 1) * rpo]
            assert rel_fro(got, ref) < 6e-2, (s, n, rel_fro(got, ref))
        opts = OS.make_adam(pn)
        for o in opts:
            o.step()
        for s in range(2):
            for name in ("res_pose1.l1", "res_angle2.l2", "angles", "upscale"):
                L = m.nets[s].layers[name]
                d_gpu = L.W.cpu() - nets[s][name + ".weight"]
                d_ref = pn[s][name + ".weight"].detach() - nets[s][name + ".weight"]
                cos = ((d_gpu * d_ref).sum() / (d_gpu.norm() * d_ref.norm())).item()
                assert cos > 0.9, (s, name, cos)
    dist.barrier()
    open(os.path.join(out_dir, "ok_%s_%s_%d" % (kind, "push_global", rank)), "w").write("ok")
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
