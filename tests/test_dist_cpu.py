"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: sharding invariants, the single final eval
reduction and gradient averaging -- the equalities the GPU path relies on."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from links_b200.shard import average_gradients, reduce_eval_sums, shard_rows
    from links_b200.synth import synth_poses, synth_pred_3d
    from oracle import metrics as OM
    from oracle import nets as ON, steps as OS
    torch.set_num_threads(2)
    # ---- eval: per-rank sums of the oracle metrics + one reduction == metrics of the whole set
    n = 101
    p2d, gt = synth_poses(n, seed=5)
    pred = synth_pred_3d(gt, seed=6)
    g, p = torch.from_numpy(gt), torch.from_numpy(pred)
    gs, ps = shard_rows(g, rank, world, multiple=1), shard_rows(p, rank, world, multiple=1)
    sums = torch.stack((OM.mpjpe(gs, ps, num_joints=17, root_joint=0).double().sum(),
                        torch.from_numpy(OM.pmpjpe_best_batch(gs.numpy(), ps.numpy())).double().sum()))
    means, total = reduce_eval_sums(sums, gs.shape[0])
    ref = (OM.mpjpe(g, p, num_joints=17, root_joint=0).double().mean().item(), float(OM.pmpjpe_best_batch(gt, pred).mean()))
    assert total == n
    assert abs(means[0] - ref[0]) < 1e-6 and abs(means[1] - ref[1]) < 1e-6
    # ---- train: mean loss over equal even shards -> averaged gradients == gradients of the global-batch mean loss
    # (per-row loss without the batch-coupled elevation statistic: that term is the documented local-shard choice)
    B = 16
    x2d, _ = synth_poses(B, seed=9)
    x = torch.from_numpy(x2d)[:, :14]
    params = ON.init_lifter_params(7, 3)
    def loss_and_grad(xb):
        pr = OS.params_require_grad(params)
        xd, xa = ON.lifter_forward(xb, pr)
        ((xd ** 2).sum(1) + xa[:, 0]).mean().backward()
        return torch.cat([pr[k].grad.flatten() for k in sorted(pr) if pr[k].grad is not None])
    g_full = loss_and_grad(x)
    xs = shard_rows(x, rank, world)
    assert xs.shape[0] % 2 == 0 and xs.shape[0] == B // world
    g_avg = average_gradients(loss_and_grad(xs), world)
    assert (g_avg - g_full).abs().max().item() < 1e-6 * max(1.0, g_full.abs().max().item())
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([1]))
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d.npy" % r)) for r in range(world))


def test_shard_bounds_properties():
    sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
    from links_b200.shard import shard_bounds
    for n, world in ((8192, 8), (1024, 4), (10, 2), (1_000_001, 8)):
        cover = []
        for r in range(world):
            b, e = shard_bounds(n, r, world)
            assert b % 2 == 0                        # row pairs never straddle ranks
            if r < world - 1:
                assert (e - b) % 2 == 0
            cover.append((b, e))
        assert cover[0][0] == 0 and cover[-1][1] == n
        assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard_bounds(3, 0, 4)
