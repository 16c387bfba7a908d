"""Sweep the SM reservation of the window chains (forward pass 2 + its dgrad chain run beside the part-flow NLL kernels) and of
the tail chain at a given batch: ms per merged step (graph replay).  python scratch/sweep_reserve.py B r_win[,r_win...] [r_tail,...]"""
import os, sys, json, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
import bench
from links_b200.steps import LifterStep

B = int(sys.argv[1])
wins = [int(v) for v in sys.argv[2].split(",")]
tails = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1]
nets, flows, full = bench.make_weights()
data = bench.make_inputs(B, 0)
for rw in wins:
    for rt in tails:
        cfg = {"prefetch_sample": True, "store_rot_2d": False}
        if rw >= 0:
            cfg["reserve_window"] = rw
        if rt >= 0:
            cfg["reserve_tail"] = rt
        step = LifterStep("both", B, nets, flows, full, cfg=cfg)
        d = {k: v.cuda() for k, v in data[0].items()}
        step.x.copy_(d["x"]); step.noise.copy_(d["noise"]); step.eps_x.copy_(d["eps_x"]); step.u_y.copy_(d["u_y"])
        step.prime()
        step.step()
        torch.cuda.synchronize()
        step.capture(warmup=1)
        for _ in range(3):
            step.graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            step.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"B": B, "reserve_window": rw, "reserve_tail": rt, "ctas_window": step._ctas_window,
                          "ctas_tail": step._ctas_tail, "ms_per_step": ms, "poses_per_s": B / ms * 1e3}), flush=True)
        del step
        gc.collect(); torch.cuda.empty_cache()
