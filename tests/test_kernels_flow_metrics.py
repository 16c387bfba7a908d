"""CPU check of the flow, metric, occlusion and elementwise CUDA kernels (csrc/*.cuh via tests/hostsim)
against the oracle and the golden fixtures generated from the reference's own utils."""
import ctypes as C

import numpy as np
import pytest
import torch

from hostsim_util import bf16_to_f32, f32_to_bf16, pack_flow, backend_params
from links_b200 import maps as MP
from links_b200.synth import synth_poses, synth_pred_3d
from oracle import flow as OF
from oracle import geometry as OG
from oracle import metrics as OM
from oracle import steps as OS


@pytest.mark.parametrize("Cdim,M", [(14, 40), (34, 33), (22, 5)])
@backend_params
def test_flow_forward_reverse_nll(Cdim, M, backend):
    L = backend
    params = OF.init_flow_params(Cdim, 50 + Cdim, perturb=0.3)
    packed = pack_flow(L, params, Cdim)
    g = torch.Generator().manual_seed(Cdim)
    x = torch.randn(M, Cdim, generator=g) * 0.2
    z_ref, ld_ref = OF.inn_forward(x, params)
    xn = np.ascontiguousarray(x.numpy())
    z = np.zeros((M, Cdim), np.float32)
    ld = np.zeros(M, np.float32)
    assert L.call("flow_apply", packed, Cdim, 8, xn, M, 0, z, ld) == 0
    np.testing.assert_allclose(z, z_ref.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ld, ld_ref.numpy(), rtol=2e-4, atol=2e-5)
    # reverse restores the input and negates the log-det
    xr = np.zeros((M, Cdim), np.float32)
    ldr = np.zeros(M, np.float32)
    assert L.call("flow_apply", packed, Cdim, 8, z, M, 1, xr, ldr) == 0
    np.testing.assert_allclose(xr, xn, rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(ldr, -ld, rtol=1e-3, atol=1e-4)
    x_ref_rev, ld_ref_rev = OF.inn_forward(z_ref, params, rev=True)
    np.testing.assert_allclose(xr, x_ref_rev.numpy(), rtol=1e-3, atol=2e-5)
    # fused NLL forward+backward vs autograd
    xg = x.clone().requires_grad_(True)
    zz, ll = OF.inn_forward(xg, params)
    nll = OF.nll(zz, ll)
    scale = 1.0 / M
    (nll.sum() * scale).backward()
    nll_sum = np.zeros(1, np.float32)
    dx = np.zeros((M, Cdim), np.float32)
    assert L.call("flow_nll_fwdbwd", packed, Cdim, 8, xn, M, scale, nll_sum, dx, None) == 0
    np.testing.assert_allclose(nll_sum[0], nll.sum().item(), rtol=1e-4)
    gref = xg.grad.numpy()
    np.testing.assert_allclose(dx, gref, rtol=2e-3, atol=2e-5 * np.abs(gref).max())
    # general VJP (autograd of the FrEIA shim): arbitrary seeds on z and log_jac_det
    gz = torch.randn(M, Cdim, generator=g) * 0.3
    gld = torch.randn(M, generator=g)
    xg = x.clone().requires_grad_(True)
    zz, ll = OF.inn_forward(xg, params)
    ((zz * gz).sum() + (ll * gld).sum()).backward()
    dx2 = np.zeros((M, Cdim), np.float32)
    assert L.call("flow_vjp", packed, Cdim, 8, xn, M, np.ascontiguousarray(gz.numpy()), np.ascontiguousarray(gld.numpy()), dx2) == 0
    gref = xg.grad.numpy()
    np.testing.assert_allclose(dx2, gref, rtol=2e-3, atol=2e-5 * np.abs(gref).max())


@backend_params
def test_flow_sample_block(backend):
    L = backend
    B = 9
    params = OF.init_flow_params(34, 40, perturb=0.3)
    packed = pack_flow(L, params, 34)
    x2d, _ = synth_poses(B, seed=3)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(B, 34, generator=g)
    ref = OS.sample_poses(torch.from_numpy(x2d), params, noise).numpy()
    out = np.zeros((2 * B, 34), np.float32)
    assert L.call("flow_sample", packed, 8, x2d, np.ascontiguousarray(noise.numpy()), B, out) == 0
    np.testing.assert_array_equal(out[:B], x2d)
    assert np.all(out[B:, 0] == 0) and np.all(out[B:, 17] == 0)
    np.testing.assert_allclose(out, ref, rtol=2e-3, atol=2e-5)


@backend_params
def test_metrics_vs_golden(golden, backend):
    L = backend
    G = golden["metrics"]
    gt, pred = np.ascontiguousarray(G["gt"]), np.ascontiguousarray(G["pred"])
    M = gt.shape[0]
    for nj, rj in ((17, 0), (16, 6)):
        g = np.ascontiguousarray(gt.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj))
        p = np.ascontiguousarray(pred.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj))
        for scaling in (1, 0):
            per = np.zeros(M, np.float32)
            mx = np.zeros(M, np.float32)
            dist = np.zeros((M, nj), np.float32)
            s = np.zeros(1, np.float64)
            assert L.call("mpjpe", g, p, M, nj, rj, scaling, per, mx, dist, s) == 0
            ref = G["mpjpe_j%d_s%d" % (nj, scaling)]
            assert np.abs(per - ref).max() < 0.05          # mm, north-star tolerance
            np.testing.assert_allclose(per, ref, rtol=2e-5)
            np.testing.assert_allclose(s[0], ref.astype(np.float64).sum(), rtol=1e-6)
            np.testing.assert_allclose(mx, dist.max(1))
        # PCK / AUC / CPS from integer threshold counts (scaling on, like the defaults)
        per = np.zeros(M, np.float32); mx = np.zeros(M, np.float32); dist = np.zeros((M, nj), np.float32)
        L.call("mpjpe", g, p, M, nj, rj, 1, per, mx, dist, None)
        thr = np.array([150.0], np.float32)
        cnt = np.zeros(1, np.uint64)
        L.call("threshold_counts", dist, dist.size, thr, 1, 1, cnt)
        np.testing.assert_allclose(cnt[0] / (M * nj) * 100, G["pck_j%d" % nj], rtol=1e-6)
        thr = torch.linspace(0, 150, 150).numpy()
        cnt = np.zeros(150, np.uint64)
        L.call("threshold_counts", dist, dist.size, thr, 150, 1, cnt)
        ref_counts = np.array([(dist < t).sum() for t in thr], np.uint64)
        np.testing.assert_array_equal(cnt, ref_counts)      # integer, bit-exact
        np.testing.assert_allclose((cnt / (M * nj * 150.0)).sum(), G["auc_j%d" % nj], rtol=1e-5)
        thr = torch.linspace(0, 300, 301).numpy()
        cnt = np.zeros(301, np.uint64)
        L.call("threshold_counts", mx, M, thr, 301, 0, cnt)
        np.testing.assert_allclose((cnt / M).sum(), G["getall_CPS_j%d" % nj], rtol=1e-5)
        # PA-MPJPE, batch semantics (metrics_batch.py:104-159)
        pa = np.zeros(M, np.float32)
        s = np.zeros(1, np.float64)
        al = np.zeros((M, 3 * nj), np.float32)
        assert L.call("pmpjpe", g, p, M, nj, 0, pa, al, s) == 0
        ref_al = OM.procrustes_batch(torch.from_numpy(p).reshape(M, 3, nj), torch.from_numpy(g).reshape(M, 3, nj))
        assert np.abs(al - ref_al.reshape(M, -1).numpy()).max() < 0.05
        ref = G["pmpjpe_batch_j%d" % nj]
        assert np.abs(pa - ref).max() < 0.05, np.abs(pa - ref).max()
    # PA-MPJPE 'best' (metrics.py:35-171, the number the scripts report; fp64 numpy reference incl. mirrored poses)
    pa = np.zeros(M, np.float32)
    assert L.call("pmpjpe", gt, pred, M, 17, 1, pa, None, None) == 0
    assert np.abs(pa - G["pmpjpe_np_best"]).max() < 0.05, np.abs(pa - G["pmpjpe_np_best"]).max()


@pytest.mark.parametrize("nj,rj", [(17, 0), (16, 6), (12, 3)])
@backend_params
def test_metrics_multichunk(nj, rj, backend):
    """Several 64-pose chunks per block + a ragged tail: exercises the staging pipeline (bulk copies on the GPU,
    cooperative copies in the host simulation) and the rotated joint order used for even row strides."""
    L = backend
    M = 64 * 7 + 13
    _, gt = synth_poses(M, seed=21)
    pred = synth_pred_3d(gt, seed=22)
    g = np.ascontiguousarray(gt.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj))
    p = np.ascontiguousarray(pred.reshape(M, 3, 17)[:, :, :nj].reshape(M, 3 * nj))
    gt_t, p_t = torch.from_numpy(g), torch.from_numpy(p)
    for scaling in (1, 0):
        per = np.zeros(M, np.float32); mx = np.zeros(M, np.float32); dist = np.zeros((M, nj), np.float32)
        s = np.zeros(1, np.float64)
        assert L.call("mpjpe", g, p, M, nj, rj, scaling, per, mx, dist, s) == 0
        ref = OM.mpjpe(gt_t, p_t, use_scaling=bool(scaling), root_joint=rj, num_joints=nj).numpy()
        assert np.abs(per - ref).max() < 0.05
        np.testing.assert_allclose(per, ref, rtol=3e-5)
        np.testing.assert_allclose(per, dist.mean(1), rtol=1e-5)
        np.testing.assert_allclose(mx, dist.max(1))
        np.testing.assert_allclose(s[0], ref.astype(np.float64).sum(), rtol=1e-6)
    pa = np.zeros(M, np.float32); al = np.zeros((M, 3 * nj), np.float32); s = np.zeros(1, np.float64)
    assert L.call("pmpjpe", g, p, M, nj, 0, pa, al, s) == 0
    ref = OM.pmpjpe_batch(gt_t, p_t, num_joints=nj).numpy()
    assert np.abs(pa - ref).max() < 0.05, np.abs(pa - ref).max()
    ref_al = OM.procrustes_batch(p_t.reshape(M, 3, nj), gt_t.reshape(M, 3, nj)).reshape(M, -1).numpy()
    assert np.abs(al - ref_al).max() < 0.05
    np.testing.assert_allclose(s[0], ref.astype(np.float64).sum(), rtol=1e-5)
    if nj == 17:
        pa1 = np.zeros(M, np.float32)
        assert L.call("pmpjpe", g, p, M, nj, 1, pa1, None, None) == 0
        assert np.abs(pa1 - OM.pmpjpe_best_batch(g, p)).max() < 0.05
        # planar predictions: rank-2 covariance -> the Newton polar iteration declines and the Jacobi SVD path (with its
        # cross-product completion) takes over; the 'best' error is still well defined
        p2 = p.copy().reshape(M, 3, nj)
        p2[::5, 2, :] = 0.0
        p2 = np.ascontiguousarray(p2.reshape(M, 3 * nj))
        assert L.call("pmpjpe", g, p2, M, nj, 1, pa1, None, None) == 0
        assert np.abs(pa1 - OM.pmpjpe_best_batch(g, p2)).max() < 0.05


@pytest.mark.parametrize("ldd", [32, 20, 17])     # 17: rows are not 16-byte multiples -> the non-bulk staging path
@backend_params
def test_eval_lift_score_fused(ldd, backend):
    L = backend
    M = 64 * 5 + 6
    p2d, gt = synth_poses(M, seed=11)
    rng = np.random.RandomState(0)
    doff = np.zeros((M, ldd), np.float32)
    doff[:, :17] = rng.normal(size=(M, 17)) * 0.3
    doff[:, 0] = 0
    pred = OG.lift(torch.from_numpy(p2d), torch.from_numpy(doff[:, :17] + 10.0)).reshape(-1, 51)
    ref = OS.eval_metrics(torch.from_numpy(gt), pred)
    pab = OM.pmpjpe_batch(torch.from_numpy(gt), pred, num_joints=17).mean().item()
    sums = np.zeros(3, np.float64)
    assert L.call("eval_lift_score", p2d, doff, ldd, gt, M, 10.0, sums) == 0
    assert abs(sums[0] / M - ref["n_mpjpe"]) < 0.05
    assert abs(sums[1] / M - ref["pa_mpjpe"]) < 0.05
    assert abs(sums[2] / M - pab) < 0.05


@backend_params
def test_index_maps_bit_exact(golden, backend):
    """pack_rows is the integer gather behind every split_* helper; compare against reference outputs."""
    L = backend
    G = golden["index_maps"]
    def gather(src, idx, period=1):
        src = np.ascontiguousarray(src.reshape(src.shape[0], -1), np.float32)
        M = src.shape[0]
        n_idx = len(idx) // period
        idx_a = np.array(idx, np.int32)
        dst = np.zeros((M, 64), np.uint16)
        dstT = np.zeros((n_idx, M + 8), np.uint16)
        assert L.call("pack_rows", src, src.shape[1], M, idx_a, n_idx, period, dst, dstT, M + 8, 2) == 0
        np.testing.assert_array_equal(dstT[:, 2:2 + M], dst[:, :n_idx].T)
        assert np.all(dst[:, n_idx:] == 0)
        return bf16_to_f32(dst[:, :n_idx])
    # values are small integers -> exactly representable in bf16? use an index-valued probe < 256
    a34 = (np.arange(6 * 34) % 251).astype(np.float32).reshape(6, 34)
    ref_l, ref_r = OG.split_data_left_right(torch.from_numpy(a34))
    np.testing.assert_array_equal(gather(a34, MP.part_index(MP.LEFT_JOINTS)), ref_l.numpy())
    np.testing.assert_array_equal(gather(a34, MP.part_index(MP.RIGHT_JOINTS)), ref_r.numpy())
    np.testing.assert_array_equal(gather(a34, MP.part_index(MP.LEG_JOINTS)), OG.part_2d(torch.from_numpy(a34), OG.LEG_JOINTS).numpy())
    # the golden fixture pins oracle == reference for these maps on an arange probe
    assert np.array_equal(OG.split_data_left_right(torch.from_numpy(G["split_lr_in"]))[0].numpy(), G["split_lr_left"])
    a51 = (np.arange(6 * 51) % 251).astype(np.float32).reshape(6, 3, 17)
    tg, inp = OS.occ_targets_inputs(torch.from_numpy(a51))
    for n in MP.OCC_NAMES:
        idx, period = MP.occ_input_index(n)
        np.testing.assert_array_equal(gather(a51, idx, period), inp[n].numpy(), err_msg=n)
        tidx = MP.occ_target_index(n)
        np.testing.assert_array_equal(a51.reshape(6, 51)[:, tidx], tg[n].numpy(), err_msg=n)
    # combine maps (helpers.py:40-53) as encoded in the geometry map struct
    for kind, choice_tabs in (("lr", (OG.COMBINE_LEFT, OG.COMBINE_RIGHT)),):
        m = MP.geom_maps(kind)
        for v, tab in enumerate(choice_tabs):
            assert [(m.src_net[v][j], m.col[j]) for j in range(17)] == tab


@backend_params
def test_occlusion_kernels(backend):
    L = backend
    M = 6
    x2d, _ = synth_poses(M, seed=2)
    rng = np.random.RandomState(4)
    hl = np.zeros((M, 32), np.float32); ht = np.zeros((M, 32), np.float32)
    hl[:, :7] = rng.normal(size=(M, 7)) * 0.5
    ht[:, :10] = rng.normal(size=(M, 10)) * 0.5
    pred = torch.cat((torch.from_numpy(hl[:, :7]), torch.from_numpy(ht[:, :10])), dim=1).clone()
    pred[:, 0] = 0
    ref_pose = OS._lift_centered(torch.from_numpy(x2d), pred, 10.0, clamp=False)
    pose = np.zeros((M, 51), np.float32)
    assert L.call("occ_lift", x2d, hl, ht, M, 10.0, pose) == 0
    np.testing.assert_allclose(pose, ref_pose.reshape(M, 51).numpy(), rtol=1e-6, atol=1e-6)
    u = rng.uniform(size=M).astype(np.float32)
    ut = torch.from_numpy(u).reshape(-1, 1)
    Ry = OG.euler_angles_to_matrix(torch.cat((torch.zeros_like(ut), (ut - 0.5) * 1.99 * np.pi, torch.zeros_like(ut)), 1), "XYZ")
    ref_rot = Ry.matmul(ref_pose).reshape(M, 51).numpy()
    rot = np.zeros((M, 51), np.float32)
    assert L.call("occ_rotate_y", pose, u, M, rot) == 0
    np.testing.assert_allclose(rot, ref_rot, rtol=1e-5, atol=1e-6)
    tidx = np.array(MP.occ_target_index("left_side"), np.int32)
    predn = np.zeros((M, 32), np.float32)
    predn[:, :18] = rng.normal(size=(M, 18))
    loss = np.zeros(1, np.float32)
    g = np.zeros((M, 64), np.uint16)
    gT = np.zeros((18, M), np.uint16)
    assert L.call("occ_mse", predn, 32, rot, tidx, 18, M, 1.0 / M, loss, g, gT, M, 0) == 0
    tgt = rot[:, tidx]
    np.testing.assert_allclose(loss[0], ((predn[:, :18] - tgt) ** 2).sum(), rtol=1e-5)
    np.testing.assert_allclose(bf16_to_f32(g)[:, :18], 2.0 / M * (predn[:, :18] - tgt), rtol=8e-3, atol=1e-6)


@backend_params
def test_adam_colsum_cast(backend):
    L = backend
    rng = np.random.RandomState(0)
    n = 1000
    p0 = rng.normal(size=n).astype(np.float32)
    tp = torch.from_numpy(p0.copy()).requires_grad_(True)
    opt = torch.optim.Adam([tp], lr=2e-4, weight_decay=1e-5)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 5):
        g = rng.normal(size=n).astype(np.float32) * 0.1
        tp.grad = torch.from_numpy(g.copy())
        opt.step()
        if step == 4:   # device-side learning rate (wins over the by-value argument): graph replays follow the scheduler
            cnt, lr_dev = np.array([step - 1], np.int32), np.array([2e-4], np.float32)
            assert L.call("adam_step", p, g, m, v, n, 123.0, 0.9, 0.999, 1e-8, 1e-5, 0, cnt, 1.0, lr_dev) == 0
        elif step < 3:
            assert L.call("adam_step", p, g, m, v, n, 2e-4, 0.9, 0.999, 1e-8, 1e-5, step, None, 1.0, None) == 0
        else:   # device-side step counter (CUDA-graph friendly): holds the number of completed steps
            cnt = np.array([step - 1], np.int32)
            assert L.call("adam_step", p, g, m, v, n, 2e-4, 0.9, 0.999, 1e-8, 1e-5, 0, cnt, 1.0, None) == 0
            assert cnt[0] == step
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=0, atol=2e-7)
    Gm = rng.normal(size=(300, 40)).astype(np.float32)
    Gb = f32_to_bf16(Gm)
    out = np.zeros(40, np.float32)
    assert L.call("colsum_bf16", Gb, 40, 300, 40, out, 0) == 0
    np.testing.assert_allclose(out, bf16_to_f32(Gb).sum(0), rtol=1e-5, atol=1e-5)
    W = rng.normal(size=(70, 22)).astype(np.float32)
    Wb = np.full((70, 64), 0x7FFF, np.uint16)
    WT = np.full((22, 72), 0x7FFF, np.uint16)
    assert L.call("cast_weight", W, 70, 22, Wb, 64, WT, 72) == 0
    np.testing.assert_array_equal(Wb[:, :22], f32_to_bf16(W))
    assert np.all(Wb[:, 22:] == 0)
    np.testing.assert_array_equal(WT[:, :70], f32_to_bf16(W).T)
    assert np.all(WT[:, 70:] == 0)


@backend_params
def test_normalize_head_kernel(backend):
    """Device normalize_head / normalize_head_test (+ transpose-flatten) vs the drop-in utils.helpers functions, which
    tests/test_dropin_cpu.py pins against the reference's outputs (golden/helpers_extra.npz)."""
    from utils import helpers as H
    L = backend
    rng = np.random.RandomState(4)
    n = 1000
    raw = (rng.normal(size=(n, 17, 2)) * 200 + 500).astype(np.float32)
    flat = np.ascontiguousarray(raw.transpose(0, 2, 1).reshape(n, 34))
    ref = H.normalize_head(flat.astype(np.float64).copy())
    out = np.zeros((n, 34), np.float32)
    acc = np.zeros(1, np.float64)
    assert L.call("normalize_head", raw, n, 0, 0, 0.0, out, acc) == 0
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=1e-7)
    out2 = np.zeros((n, 34), np.float32)
    assert L.call("normalize_head", flat, n, 0, 1, 0.0, out2, acc) == 0           # already flattened rows
    np.testing.assert_allclose(out2, ref, rtol=2e-5, atol=1e-7)
    ref_t = H.normalize_head_test(flat.astype(np.float64).copy())
    assert L.call("normalize_head", raw, n, 0, 0, 145.40964, out, None) == 0
    np.testing.assert_allclose(out, ref_t, rtol=2e-5, atol=1e-7)
