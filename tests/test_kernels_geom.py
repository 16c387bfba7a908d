"""CPU check of the geometry/loss CUDA kernels (csrc/geom.cuh via tests/hostsim) against the oracle.

The kernel math (forward, losses and the hand-derived backward incl. the batch-statistic coupling through
props.mean()/std()) is compared with torch autograd over the oracle restatement of
train_leg_torso_lifter.py:153-272 / train_left_right_lifter.py:150-423 in fp64.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from hostsim_util import View, bf16_to_f32, backend_params
from links_b200 import maps as MP
from oracle import geometry as OG
from oracle import steps as OS

HEAD_LD = 32


def make_inputs(kind, N, seed, clamp=False, root_zero=True):
    rng = np.random.RandomState(seed)
    nj = (7, 10) if kind == "lt" else (11, 11)
    u = (rng.normal(size=(N, 34)) * 0.12).astype(np.float32)
    if root_zero:   # what normalize_head / the flow sampler produce; the kernels must not depend on it
        u[:, 0] = 0.0
        u[:, 17] = 0.0
    heads = [np.zeros((N, HEAD_LD), np.float32) for _ in range(2)]
    heads2 = [np.zeros((N, HEAD_LD), np.float32) for _ in range(2)]
    angs = [np.zeros((N, HEAD_LD), np.float32) for _ in range(2)]
    for p in range(2):
        heads[p][:, :nj[p]] = rng.normal(size=(N, nj[p])) * 0.6
        heads2[p][:, :nj[p]] = rng.normal(size=(N, nj[p])) * 0.6
        angs[p][:, 0] = rng.normal(size=N) * 0.3 + 0.2
    if clamp:  # force the depth < 1 branch (:186, :232) on some joints
        heads[0][::3, 2] = -9.5
        heads[1][1::4, 3] = -12.0
        heads2[0][::2, 1] = -9.7
    eps = rng.normal(size=N).astype(np.float32)
    uy = rng.uniform(size=N).astype(np.float32)
    ext = [rng.normal(size=(N, 2 * nj[p])).astype(np.float32) * 0.05 for p in range(2)]
    ext_l = [np.zeros((N, HEAD_LD), np.float32) for _ in range(2)]
    for p in range(2):
        ext_l[p][:, :2 * nj[p]] = rng.normal(size=(N, 2 * nj[p])) * 0.05
    return dict(u=u, heads=heads, heads2=heads2, angs=angs, eps=eps, uy=uy, ext=ext, ext_l=ext_l, nj=nj)


def oracle_eval(kind, inp, cfg):
    """fp64 oracle: losses, qparts, and autograd gradients w.r.t. heads / heads2 / angle heads."""
    dt = torch.float64
    nj = inp["nj"]
    u = torch.tensor(inp["u"], dtype=dt)
    H = [torch.tensor(inp["heads"][p][:, :nj[p]], dtype=dt, requires_grad=True) for p in range(2)]
    H2 = [torch.tensor(inp["heads2"][p][:, :nj[p]], dtype=dt, requires_grad=True) for p in range(2)]
    A = [torch.tensor(inp["angs"][p][:, :1], dtype=dt, requires_grad=True) for p in range(2)]
    eps, uy = torch.tensor(inp["eps"], dtype=dt), torch.tensor(inp["uy"], dtype=dt)
    depth = cfg["depth"]
    zero0 = lambda t: torch.cat((torch.zeros_like(t[:, :1]), t[:, 1:]), dim=1)
    props = (A[0] + A[1]) / 2
    R = OS._rotation(props, eps, uy)
    if kind == "lt":
        bone = torch.tensor(OG.BONE_REL_MPI, dtype=dt)
        preds = [zero0(torch.cat((H[0], H[1]), dim=1))]
        preds2 = [zero0(torch.cat((H2[0], H2[1]), dim=1))]
        joints = (OG.LEG_JOINTS, OG.TORSO_JOINTS)
        part_variant = (0, 0)
    else:
        bone = torch.tensor(OG.BONE_REL_H36M, dtype=dt)
        preds = [zero0(OG.combine_left_right_1d(H[0], H[1], c)) for c in ("left", "right")]
        preds2 = [zero0(OG.combine_left_right_1d(H2[0], H2[1], c)) for c in ("left", "right")]
        joints = (OG.LEFT_JOINTS, OG.RIGHT_JOINTS)
        part_variant = (0, 1)
    terms = [0.0, 0.0, 0.0, 0.0]
    q = []
    for v in range(len(preds)):
        p3 = OS._lift_centered(u, preds[v], depth)
        rot, rot2d = OS._rotate_project(R, p3, depth)
        q.append(rot2d)
        t = OS._consistency_terms(u, R, p3, rot, rot2d, preds2[v], bone, depth)
        for i in range(4):
            terms[i] = terms[i] + t[i]
    L3d, rep, pair, bl = terms
    qparts = [OG.part_2d(q[part_variant[p]], joints[p]) for p in range(2)]
    loss = cfg["weight_3d"] * L3d + cfg["weight_2d"] * rep + cfg["weight_velocity"] * pair + cfg["weight_bl"] * bl
    for p in range(2):
        e = torch.tensor(inp["ext"][p], dtype=dt) + torch.tensor(inp["ext_l"][p][:, :2 * nj[p]], dtype=dt)
        loss = loss + (e * qparts[p]).sum()
    loss.backward()
    return dict(L3d=L3d.item(), rep=rep.item(), pair=pair.item(), bl=bl.item(), qparts=[t.detach().numpy() for t in qparts],
                q=[t.detach().numpy() for t in q], dH=[h.grad.numpy() for h in H], dH2=[h.grad.numpy() for h in H2],
                dA=[a.grad.numpy() for a in A], stats=(props.mean().item(), props.std().item()))


@pytest.mark.parametrize("kind,N,clamp,root_zero", [("lt", 10, False, True), ("lr", 10, False, True), ("lt", 7, True, True),
                                                     ("lr", 6, True, True), ("lt", 21, False, False),
                                                     ("lr", 19, True, False), ("lt", 3, False, False)])
@backend_params
def test_geometry_kernels_match_oracle(kind, N, clamp, root_zero, backend):
    L = backend
    cfg = dict(OS.DEFAULT_CFG)
    inp = make_inputs(kind, N, seed=5 + N, clamp=clamp, root_zero=root_zero)
    ref = oracle_eval(kind, inp, cfg)
    m = MP.geom_maps(kind, cfg)
    nj = inp["nj"]
    stats = np.zeros(2, np.float32)
    assert L.call("elev_stats", inp["angs"][0], inp["angs"][1], N, stats) == 0
    np.testing.assert_allclose(stats, ref["stats"], rtol=2e-5)

    # ---- forward
    qp = [np.zeros((N, 2 * nj[p]), np.float32) for p in range(2)]
    qf = [np.zeros((N, 34), np.float32) for _ in range(2)]
    common = [inp["u"], inp["heads"][0], inp["heads"][1], inp["angs"][0], inp["angs"][1],
              inp["eps"], inp["uy"], stats]
    assert L.call("geom_forward", C.byref(m), *common, N, qp[0], qp[1], qf[0], qf[1]) == 0
    for p in range(2):
        np.testing.assert_allclose(qp[p], ref["qparts"][p], rtol=2e-4, atol=2e-6)
    for v in range(m.V):
        np.testing.assert_allclose(qf[v], ref["q"][v], rtol=2e-4, atol=2e-6)

    # ---- losses + d/d(pass-2 heads)
    ldT = ((N + 7) // 8) * 8 + 8
    sums = np.zeros(4, np.float32)
    g2 = [np.zeros((N, 64), np.uint16) for _ in range(2)]
    g2T = [np.zeros((nj[p], ldT), np.uint16) for p in range(2)]
    assert L.call("geom_loss", C.byref(m), *common, inp["heads2"][0], inp["heads2"][1], N, sums,
                           g2[0], g2[1], g2T[0], g2T[1], ldT, 3) == 0
    npairs = N // 2
    np.testing.assert_allclose(sums[0] / N, ref["L3d"], rtol=3e-5)
    np.testing.assert_allclose(sums[1] / N, ref["rep"], rtol=3e-5)
    np.testing.assert_allclose(sums[2] / npairs, ref["pair"], rtol=3e-5)
    np.testing.assert_allclose(sums[3] / N, ref["bl"], rtol=3e-5)
    for p in range(2):
        got = bf16_to_f32(g2[p])[:, :nj[p]]
        scale = np.abs(ref["dH2"][p]).max()
        np.testing.assert_allclose(got, ref["dH2"][p], rtol=8e-3, atol=2e-5 * scale)
        assert np.all(bf16_to_f32(g2[p])[:, nj[p]:] == 0)
        np.testing.assert_array_equal(g2T[p][:, 3:3 + N], g2[p][:, :nj[p]].T)   # transposed copy, bit-exact

    # ---- full backward
    g1 = [np.zeros((N, 64), np.uint16) for _ in range(2)]
    g1T = [np.zeros((nj[p], ldT), np.uint16) for p in range(2)]
    dgam, da, red = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(2, np.float32)
    assert L.call("geom_backward", C.byref(m), *common, inp["heads2"][0], inp["heads2"][1],
                               inp["ext"][0], inp["ext"][1], inp["ext_l"][0], inp["ext_l"][1], N,
                               g1[0], g1[1], g1T[0], g1T[1], ldT, 0, dgam, da, red) == 0
    for p in range(2):
        got = bf16_to_f32(g1[p])[:, :nj[p]]
        scale = np.abs(ref["dH"][p]).max()
        np.testing.assert_allclose(got, ref["dH"][p], rtol=8e-3, atol=2e-5 * scale)
        np.testing.assert_array_equal(g1T[p][:, :N], g1[p][:, :nj[p]].T)
    np.testing.assert_allclose(red[0], da.sum(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(red[1], (da * inp["eps"]).sum(), rtol=1e-4, atol=1e-6)
    ga = [np.zeros((N, 64), np.uint16) for _ in range(2)]
    gaT = [np.zeros((1, ldT), np.uint16) for _ in range(2)]
    assert L.call("geom_backward_angles", inp["angs"][0], inp["angs"][1], inp["eps"], stats, dgam,
                                      red, N, ga[0], ga[1], gaT[0], gaT[1], ldT, 0, 0) == 0
    for p in range(2):
        got = bf16_to_f32(ga[p])[:, 0]
        scale = np.abs(ref["dA"][p]).max()
        np.testing.assert_allclose(got, ref["dA"][p][:, 0], rtol=8e-3, atol=5e-5 * scale)


@pytest.mark.parametrize("kind,N", [("lt", 37), ("lr", 64)])
@backend_params
def test_packed_head_rows_equal_separate_arrays(kind, N, backend):
    """The staging plan groups strided tensors whose used columns share one 128-byte window (the packed head rows of
    MlpSet(head_groups=...)): forward / loss / backward on column-offset views into ONE [N, 32] buffer per pass must equal,
    bit for bit, the same calls on four separate arrays (one staged region instead of four; same arithmetic)."""
    L = backend
    cfg = dict(OS.DEFAULT_CFG)
    inp = make_inputs(kind, N, seed=91 + N)
    m = MP.geom_maps(kind, cfg)
    nj = inp["nj"]
    stats = np.zeros(2, np.float32)
    assert L.call("elev_stats", inp["angs"][0], inp["angs"][1], N, stats) == 0
    c_ang = (nj[0] + nj[1], nj[0] + nj[1] + 1)
    pack1, pack2 = np.zeros((N, HEAD_LD), np.float32), np.zeros((N, HEAD_LD), np.float32)
    for p, c0 in enumerate((0, nj[0])):
        pack1[:, c0:c0 + nj[p]] = inp["heads"][p][:, :nj[p]]
        pack2[:, c0:c0 + nj[p]] = inp["heads2"][p][:, :nj[p]]
        pack1[:, c_ang[p]] = inp["angs"][p][:, 0]

    def run(packed):
        if packed:
            heads = [View(pack1, 0), View(pack1, nj[0])]
            angs = [View(pack1, c_ang[0]), View(pack1, c_ang[1])]
            heads2 = [View(pack2, 0), View(pack2, nj[0])]
        else:
            heads, angs, heads2 = inp["heads"], inp["angs"], inp["heads2"]
        common = [inp["u"], heads[0], heads[1], angs[0], angs[1], inp["eps"], inp["uy"], stats]
        qp = [np.zeros((N, 2 * nj[p]), np.float32) for p in range(2)]
        assert L.call("geom_forward", C.byref(m), *common, N, qp[0], qp[1], None, None) == 0
        sums = np.zeros(4, np.float32)
        g2 = [np.zeros((N, 64), np.uint16) for _ in range(2)]
        assert L.call("geom_loss", C.byref(m), *common, heads2[0], heads2[1], N, sums, g2[0], g2[1], None, None, 0, 0) == 0
        g1 = [np.zeros((N, 64), np.uint16) for _ in range(2)]
        dgam, da, red = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(2, np.float32)
        assert L.call("geom_backward", C.byref(m), *common, heads2[0], heads2[1], inp["ext"][0], inp["ext"][1],
                      inp["ext_l"][0], inp["ext_l"][1], N, g1[0], g1[1], None, None, 0, 0, dgam, da, red) == 0
        return qp + g2 + g1 + [dgam, da]

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@backend_params
def test_unaligned_dense_tensor_is_refused(backend):
    """Dense tensors are staged in 16-byte chunks: a pose buffer that is not 16-byte aligned is an error, not a fault."""
    L = backend
    N = 16
    cfg = dict(OS.DEFAULT_CFG)
    inp = make_inputs("lt", N, seed=3)
    m = MP.geom_maps("lt", cfg)
    stats = np.array([0.1, 0.2], np.float32)
    big = np.zeros(N * 34 + 4, np.float32)
    qp = [np.zeros((N, 14), np.float32), np.zeros((N, 20), np.float32)]
    rc = L.call("geom_forward", C.byref(m), View(big, 1), inp["heads"][0], inp["heads"][1], inp["angs"][0], inp["angs"][1],
                inp["eps"], inp["uy"], stats, N, qp[0], qp[1], None, None)
    assert rc == -2          # LINKS_E_ALIGN
