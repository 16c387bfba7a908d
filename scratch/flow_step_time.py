"""ms per flow training step (config #1: full-pose flow B = 256; the four part flows B = 256), eager launches vs graph replay."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
from links_b200 import init as INIT
from links_b200.flowtrain import FlowTrainStep, PartFlowTrainer
from links_b200.synth import synth_poses

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x2d, _ = synth_poses(B, seed=77)


def timed(fn, n=30):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name in ("full", "parts"):
    for mode in ("eager", "graph"):
        if name == "full":
            t = FlowTrainStep(34, INIT.init_flow_params(34, 40), B, lr=2e-4)
        else:
            width = {"legs": 14, "torso": 20, "left": 22, "right": 22}
            t = PartFlowTrainer(INIT.init_flow_params(34, 40), {n: INIT.init_flow_params(width[n], 50 + i) for i, n in enumerate(PartFlowTrainer.NAMES)}, B)
        t.x.copy_(torch.from_numpy(x2d)); t.noise.normal_()
        ms = timed(t.step if mode == "eager" else t.run)
        print(json.dumps({"trainer": name, "B": B, "mode": mode, "ms_per_step": ms, "poses_per_s": B / ms * 1e3, "loss": t.loss_dict()["loss"]}), flush=True)
