import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")]
import torch
from links_b200.steps import LifterStep
from links_b200.synth import synth_poses
from oracle import flow as OF, nets as ON
B = 192
nets = [ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)]
flows = [OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3)]
full = OF.init_flow_params(34, 40, perturb=0.3)
x2d, _ = synth_poses(B, seed=3)
g = torch.Generator().manual_seed(8)
d = dict(x=torch.from_numpy(x2d), noise=torch.randn(B, 34, generator=g), eps_x=torch.randn(2 * B, generator=g), u_y=torch.rand(2 * B, generator=g))
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = [LifterStep("lt", B, nets, flows, full, cfg={"fuse_adam": f}) for f in (True, False, False)]
for st in steps:
    for _ in range(nsteps):
        st.x.copy_(d["x"]); st.noise.copy_(d["noise"]); st.eps_x.copy_(d["eps_x"]); st.u_y.copy_(d["u_y"])
        st.step()
torch.cuda.synchronize()
a, b, c = steps[0].mlp, steps[1].mlp, steps[2].mlp
print("unfused vs unfused (run-to-run noise): master %.3e m %.3e v %.3e" % ((b.master - c.master).abs().max().item(), (b.exp_avg - c.exp_avg).abs().max().item(), (b.exp_avg_sq - c.exp_avg_sq).abs().max().item()))
for s in range(2):
    for n in a.layer_names:
        La, Lb = a.nets[s].layers[n], b.nets[s].layers[n]
        oa = La.off_W
        dm = (a.exp_avg[oa:oa + La.N * La.K] - b.exp_avg[oa:oa + La.N * La.K]).abs().max().item()
        mm = b.exp_avg[oa:oa + La.N * La.K].abs().max().item()
        dW = (La.W - Lb.W).abs().max().item()
        db = (La.b - Lb.b).abs().max().item()
        dmb = (a.exp_avg[La.off_b:La.off_b + La.N] - b.exp_avg[La.off_b:La.off_b + La.N]).abs().max().item()
        print("net %d %-14s fused_ok %d  dW %.3e  dm %.3e (|m| %.3e)  db %.3e dmb %.3e" % (s, n, La.fused_ok, dW, dm, mm, db, dmb))
