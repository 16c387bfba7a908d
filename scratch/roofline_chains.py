"""The four chain launches of one merged LT+LR step as pure GEMM work on all SMs (what bench.py's `roofline` times),
run REPS times: `ncu -k regex:gemm_kernel --launch-skip 4*(REPS-1) --launch-count 4` captures the last repetition."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")]
import torch
from links_b200 import init as INIT
from links_b200.mlp import MlpSet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 2 * B
nj = [7, 10, 11, 11]
m = MlpSet("lifter", [2 * n for n in nj], [{"downscale": n, "angles": 1} for n in nj], N, n_passes=2, train=True,
           pass_branches=[["pose", "angle"], ["pose"]])
m.load_state_dicts([INIT.init_lifter_params(n, 11 + i) for i, n in enumerate(nj)])
for p in range(2):
    for s in range(4):
        m.x0[p][s].normal_(0, 0.2)
        for h in ("downscale", "angles"):
            m.G[p][s][h].normal_(0, 0.05)
plans = [m._chained(("r", 0), lambda: m._build_forward(0)), m._chained(("r", 1), lambda: m._build_forward(1)),
         m._chained(("r", 2), lambda: m._build_backward(1, True)), m._chained(("r", 3), lambda: m._build_backward(0, False, None, True, False))]
ops = [op for plan in plans for op in plan if hasattr(op, "plan")]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    if r == reps - 1:
        torch.cuda.synchronize(); e0.record()
    for op in ops:
        op()
e1.record(); torch.cuda.synchronize()
flops = 2 * 2 * B * sum(2 * n * 1024 + 14 * 1024 * 1024 + 1024 * n + 1024 + 2 * n * 1024 + 8 * 1024 * 1024 + 1024 * n for n in nj) * 3
print("B %d: last repetition %.1f us, ~%.0f TFLOP/s (fwd+dgrad+wgrad, approx)" % (B, e0.elapsed_time(e1) * 1e3, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12))
