#!/usr/bin/env python
"""Isolated roofline measurements of the HBM-bound kernels (geometry, metrics, Adam) at sizes far above the 126 MB L2
(SURVEY 8d "small-working-set caveat": inside the B=1024 step these kernels move ~25 MB and are latency-bound; the
>= 70 % of HBM target is measured here at N >= 4 M rows).  achieved = ALGORITHMIC bytes / CUDA-event time; peak =
MEASURED_PEAKS.json hbm_gbs.  Prints one JSON object; `python bench_kernels.py > profiles/...json`.
Also times the sharded-eval path (config #5): lift (pose branch) + N-MPJPE + both PA-MPJPE, poses/s on one GPU."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from links_b200 import _cabi, maps  # noqa: E402
from links_b200.init import init_lifter_params  # noqa: E402
from links_b200.occlusion import EvalRunner  # noqa: E402


def timed(fn, reps=10, warm=3):
    reps = int(os.environ.get("LINKS_BK_REPS", reps))       # 1 / 0 for an ncu capture of one launch per kernel
    warm = int(os.environ.get("LINKS_BK_WARM", warm))
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    L = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    peak = 6543.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    out = {"peak_hbm_gbs": peak, "kernels": {}}

    def rec(name, bytes_per_unit, units, sec, note=""):
        gbs = bytes_per_unit * units / sec / 1e9
        out["kernels"][name] = {"units": units, "algorithmic_bytes_per_unit": bytes_per_unit, "us": sec * 1e6,
                                "achieved_gbs": gbs, "frac_of_measured_peak": gbs / peak, "note": note}

    # ---------------- geometry (LT maps), N rows
    N = 4 * 1024 * 1024
    f32 = dict(dtype=torch.float32, device="cuda")
    m = maps.geom_maps("lt")
    u = torch.randn(N, 34, **f32) * 0.1
    # all depth / angle heads of a pass side by side in ONE [N, 32] row (what LifterStep allocates: MlpSet head_groups):
    # leg depths cols 0-6, torso depths 7-16, the two angle heads 17 and 18
    pack1, pack2 = torch.randn(N, 32, **f32) * 0.1, torch.randn(N, 32, **f32) * 0.1
    heads = [pack1[:, 0:], pack1[:, 7:]]
    angs = [pack1[:, 17:], pack1[:, 18:]]
    heads2 = [pack2[:, 0:], pack2[:, 7:]]
    eps, uy = torch.randn(N, **f32), torch.rand(N, **f32)
    stats = torch.zeros(2, **f32)
    qp = [torch.zeros(N, 14, **f32), torch.zeros(N, 20, **f32)]
    qf = [torch.zeros(N, 34, **f32) for _ in range(2)]
    common = [u.data_ptr(), heads[0].data_ptr(), heads[1].data_ptr(), angs[0].data_ptr(), angs[1].data_ptr(), eps.data_ptr(),
              uy.data_ptr(), stats.data_ptr()]
    _cabi.check(L.links_elev_stats(angs[0].data_ptr(), angs[1].data_ptr(), N, stats.data_ptr(), st), "stats")
    t = timed(lambda: L.links_geom_forward(C.byref(m), *common, N, qp[0].data_ptr(), qp[1].data_ptr(), None, None, st))
    rec("geom_forward", 352, N, t, "u 34 + depth heads 17 + angles 2 + draws 2 read, projected part inputs 34 written (fp32)")
    sums = torch.zeros(4, **f32)
    g2 = [torch.zeros(N, 64, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    t = timed(lambda: L.links_geom_loss(C.byref(m), *common, heads2[0].data_ptr(), heads2[1].data_ptr(), N, sums.data_ptr(),
                                        g2[0].data_ptr(), g2[1].data_ptr(), None, None, 0, 0, st))
    rec("geom_lossgrad<0>", 624, N, t, "inputs 34+17+3+34+17 floats, outputs 17+34 (SURVEY 8d)")
    dfl = [torch.randn(N, 14, **f32), torch.randn(N, 20, **f32)]
    dli = [torch.randn(N, 32, **f32) for _ in range(2)]
    g1 = [torch.zeros(N, 64, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    dgam, da, red = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(2, **f32)
    t = timed(lambda: L.links_geom_backward(C.byref(m), *common, heads2[0].data_ptr(), heads2[1].data_ptr(), dfl[0].data_ptr(),
                                            dfl[1].data_ptr(), dli[0].data_ptr(), dli[1].data_ptr(), N, g1[0].data_ptr(),
                                            g1[1].data_ptr(), None, None, 0, 0, dgam.data_ptr(), da.data_ptr(), red.data_ptr(), st))
    rec("geom_lossgrad<1> (backward)", 492, N, t, "inputs 54+34+17 floats (+recompute), outputs 17+1 (SURVEY 8d)")
    del u, heads, angs, heads2, pack1, pack2, qp, qf, g2, g1, dfl, dli
    torch.cuda.empty_cache()

    # ---------------- metrics, M poses
    M = 8 * 1024 * 1024
    gt = torch.randn(M, 51, **f32) * 300
    pr = gt + torch.randn(M, 51, **f32) * 30
    per = torch.zeros(M, **f32)
    dsum = torch.zeros(1, dtype=torch.float64, device="cuda")
    t = timed(lambda: L.links_mpjpe(gt.data_ptr(), pr.data_ptr(), M, 17, 0, 1, per.data_ptr(), None, None, dsum.data_ptr(), st))
    rec("mpjpe (J=17, scaled)", 2 * 51 * 4 + 4, M, t)
    for mode, nm in ((0, "pmpjpe batch semantics"), (1, "pmpjpe 'best' semantics")):
        t = timed(lambda: L.links_pmpjpe(gt.data_ptr(), pr.data_ptr(), M, 17, mode, per.data_ptr(), None, dsum.data_ptr(), st))
        rec(nm, 2 * 51 * 4 + 4, M, t, "3x3 Jacobi SVD in registers")
    p2d = torch.randn(M, 34, **f32) * 0.1
    doff = torch.randn(M, 32, **f32) * 0.1
    s3 = torch.zeros(3, dtype=torch.float64, device="cuda")
    t = timed(lambda: L.links_eval_lift_score(p2d.data_ptr(), doff.data_ptr(), 32, gt.data_ptr(), M, 10.0, s3.data_ptr(), st))
    rec("eval_lift_score (lift + N-MPJPE + 2x PA-MPJPE)", 408, M, t, "2D 34 + depth 17 + GT 51 floats per pose, sums reduced in-kernel")
    del gt, pr, per, p2d, doff
    torch.cuda.empty_cache()

    # ---------------- Adam, n params
    n = 64 * 1024 * 1024
    p, g, m1, m2 = (torch.randn(n, **f32) * 0.01 for _ in range(4))
    m2.abs_()
    step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    t = timed(lambda: L.links_adam_step(p.data_ptr(), g.data_ptr(), m1.data_ptr(), m2.data_ptr(), n, 2e-4, 0.9, 0.999, 1e-8,
                                        1e-5, 0, step_dev.data_ptr(), 1.0, None, st))
    rec("adam_kernel", 28, n, t, "16 B read + 12 B written per parameter (the bf16 shadow refresh is a separate launch)")
    del p, g, m1, m2
    torch.cuda.empty_cache()

    # ---------------- config #5: sharded eval on one GPU (10 M poses / 8 GPUs = 1.25 M per GPU)
    n_eval, chunk = 1_250_000, 65536
    ev = EvalRunner("lr", [init_lifter_params(11, 13), init_lifter_params(11, 14)], chunk=chunk)
    x = torch.randn(n_eval, 34, **f32) * 0.1
    g3 = torch.randn(n_eval, 51, **f32) * 300

    def run_eval():
        ev.reset()
        for i in range(0, n_eval, chunk):
            ev.run_chunk(x[i:i + chunk], g3[i:i + chunk])
    t = timed(run_eval, reps=3, warm=1)
    out["eval_config5"] = {"poses_per_gpu": n_eval, "seconds": t, "poses_per_sec_per_gpu": n_eval / t,
                           "note": "left/right lifters (pose branch, 16.8 M MACs/pose) + fused scoring, inputs resident in HBM"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
