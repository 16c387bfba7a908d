"""ncu target: the four chain launches of one merged lifter step (forward pass 1, forward pass 2, pass-2 dgrad, pass-1 dgrad +
weight gradients + fused Adam) at batch B on all SMs, run twice (capture the second round: -k regex:gemm_kernel -s 4 -c 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
from links_b200.mlp import MlpSet
from links_b200 import init as INIT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
PURE = len(sys.argv) > 2 and sys.argv[2] == "pure"      # weight gradients stored (no fused optimiser): bench.py's `roofline` launches
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
m.adam_prepare()
cases = [("fwd0", lambda: m._chained(("c", 0), lambda: m._build_forward(0))),
         ("fwd1", lambda: m._chained(("c", 1), lambda: m._build_forward(1))),
         ("bwd1", lambda: m._chained(("c", 2), lambda: m._build_backward(1, True))),
         ("bwd0+wgrad", lambda: m._chained(("c", 5), lambda: m._build_backward(0, False, None, True, False))) if PURE else
         ("bwd0+wgrad+adam", lambda: m._chained(("c", 6), lambda: m._build_backward(0, False, None, True, True)))]
ops = [[op for op in chain() if hasattr(op, "plan")] for _, chain in cases]
for rnd in range(2):
    for (name, _), cops in zip(cases, ops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for op in cops:
            op()
        e1.record()
        torch.cuda.synchronize()
        if rnd == 1:
            print(name, "launches", len(cops), "us", round(e0.elapsed_time(e1) * 1e3, 1), "tiles", sum(o.plan.total_tiles for o in cops), flush=True)
