import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'links-3d-human-pose-estimation_b200')
from links_b200.mlp import MlpSet, WIDTH
from links_b200._cabi import EPI_LEAKY_PRE
from links_b200.init import init_lifter_params

M = 2048
mlp = MlpSet("lifter", [14, 20], [{"downscale": 7, "angles": 1}, {"downscale": 10, "angles": 1}], M, n_passes=2,
             pass_branches=[["pose", "angle"], ["pose"]])
mlp.load_state_dicts([init_lifter_params(7, 11), init_lifter_params(10, 12)])
for p in range(2):
    for s in range(2):
        mlp.x0[p][s].normal_(0, 0.1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("ldT", mlp.ldT, "pass_stride", mlp.pass_stride)
for name, plan in (("fwd0", mlp.forward_plan(0)), ("fwd1", mlp.forward_plan(1))):
    mlp.run(plan)
    print(name, [round(timeit(op), 1) for op in plan])
# dissect one pass-1 branch l1 launch: res_pose1.l1 / res_angle1.l1 of both nets
act, nets = mlp.act[0], mlp.nets


def l1_prob(s, blk, xin, **over):
    kw = dict(flags=EPI_LEAKY_PRE, bias=nets[s].layers[blk + ".l1"].b, out=act[s][blk + ".a1"], outT=mlp.actT[s][blk + ".a1"], outT_col0=0)
    kw.update(over)
    return mlp._prob(act[s][xin], nets[s].layers[blk + ".l1"].Wb, M, WIDTH, WIDTH, WIDTH, WIDTH, **kw)


cases = {
    "pose only (2 probs)": [l1_prob(s, "res_pose1", "res_common.y") for s in range(2)],
    "angle only (2 probs)": [l1_prob(s, "res_angle1", "res_common.y") for s in range(2)],
    "angle, no outT": [l1_prob(s, "res_angle1", "res_common.y", outT=None) for s in range(2)],
    "angle, no out": [l1_prob(s, "res_angle1", "res_common.y", out=None) for s in range(2)],
    "angle -> pose buffers": [l1_prob(s, "res_angle1", "res_common.y", out=act[s]["res_pose1.a1"], outT=mlp.actT[s]["res_pose1.a1"]) for s in range(2)],
    "pose+angle (4 probs)": [l1_prob(s, b, "res_common.y") for b in ("res_pose1", "res_angle1") for s in range(2)],
    "pose x2 distinct out (4 probs)": [l1_prob(s, b, "res_common.y") for b in ("res_pose1", "res_pose2") for s in range(2)],
}
for k, probs in cases.items():
    print("%-34s %.1f us" % (k, timeit(mlp._launch(probs))))
for nm in ("res_pose1.a1", "res_angle1.a1", "res_angle3.y"):
    t = mlp.actT[0][nm]
    print(nm, "ptr %x" % t.data_ptr(), "align2MB", t.data_ptr() % (2 << 20))
