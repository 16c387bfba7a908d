"""A/B of the optimiser placement in the single-GPU merged step (graph replay, ms per step):
  fused      : Adam inside the weight-gradient epilogues of the tail chain (default)
  unfused-1  : tail chain stores fp32 gradients; ONE Adam launch over the flat buffer + ONE batched shadow cast
  unfused-n  : per-level buckets (one Adam + one cast launch per bucket, issued as the buckets of the chain complete)
python scratch/ab_fuse_adam.py [B]"""
import os, sys, json, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
import bench
from links_b200.steps import LifterStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nets, flows, full = bench.make_weights()
data = bench.make_inputs(B, 0)
variants = [("fused", {}),
            ("unfused-1", {"fuse_adam": False, "dp_layout": True, "dp_buckets": 1}),
            ("unfused-n", {"fuse_adam": False}),
            ("fused", {})]
for name, extra in variants:
    cfg = {"prefetch_sample": True, "store_rot_2d": False}
    cfg.update(extra)
    step = LifterStep("both", B, nets, flows, full, cfg=cfg)
    d = {k: v.cuda() for k, v in data[0].items()}
    step.x.copy_(d["x"]); step.noise.copy_(d["noise"]); step.eps_x.copy_(d["eps_x"]); step.u_y.copy_(d["u_y"])
    step.prime()
    step.step()
    torch.cuda.synchronize()
    first = step.loss_dict()
    step.capture(warmup=1)
    for _ in range(3):
        step.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        step.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    last = step.loss_dict()
    print(json.dumps({"variant": name, "B": B, "buckets": len(step.mlp.buckets), "ctas_tail": step._ctas_tail,
                      "ms_per_step": ms, "poses_per_s": B / ms * 1e3,
                      "first_loss": {k: v["loss"] for k, v in first.items()},
                      "last_loss": {k: v["loss"] for k, v in last.items()}}),
          flush=True)
    del step
    gc.collect(); torch.cuda.empty_cache()
