"""Three eager merged (LT+LR) steps at B poses; used under ncu for the per-launch duration list of ONE step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
import bench
from links_b200.steps import LifterStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nets, flows, full = bench.make_weights()
step = LifterStep("both", B, nets, flows, full, cfg={"prefetch_sample": True})
d = bench.make_inputs(B, 0, 1)[0]
step.x.copy_(d["x"]); step.noise.copy_(d["noise"]); step.eps_x.copy_(d["eps_x"]); step.u_y.copy_(d["u_y"])
step.prime()
for _ in range(n_steps):
    step.step()
torch.cuda.synchronize()
print("launches so far:", step.lib.links_launch_count())
