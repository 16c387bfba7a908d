"""Per-chain timings of the merged (4-network) lifter engine: chain launch vs the same layers as grouped launches."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
from links_b200.mlp import MlpSet
from links_b200 import init as INIT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0
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


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


mc = ctas if ctas > 0 else None
cases = [("fwd0", lambda: m._chained(("c", 0), lambda: m._build_forward(0), max_ctas=mc), lambda: m.forward_plan(0)),
         ("fwd1", lambda: m._chained(("c", 1), lambda: m._build_forward(1), max_ctas=mc), lambda: m.forward_plan(1)),
         ("bwd1", lambda: m._chained(("c", 2), lambda: m._build_backward(1, True), max_ctas=mc), lambda: m.backward_plan(1, True)),
         ("bwd0", lambda: m._chained(("c", 3), lambda: m._build_backward(0, False), max_ctas=mc), lambda: m.backward_plan(0, False)),
         ("bwd0+wgrad", lambda: m._chained(("c", 4), lambda: m._build_backward(0, False, None, True), max_ctas=mc),
          lambda: m.backward_plan(0, False, None, True)),
         ("wgrad", lambda: m._chained(("c", 5), lambda: m._build_wgrad(), max_ctas=mc), lambda: m.wgrad_plan()),
         ("bwd0+wgrad+adam", lambda: m._chained(("c", 6), lambda: m._build_backward(0, False, None, True, True), max_ctas=mc),
          lambda: m.backward_plan(0, False, None, True, True))]
m.adam_prepare()
only = os.environ.get("CHAIN_PROF_ONLY")
if only:
    cases = [c for c in cases if c[0] in only.split(",")]
out = {}
for name, chain, grouped in cases:
    cops = [op for op in chain() if hasattr(op, "plan")]
    gops = [op for op in grouped() if not isinstance(op, tuple) and getattr(op, "__name__", "") == "run"]
    tc = timed(lambda: [op() for op in cops])
    tg = timed(lambda: [op() for op in gops])
    pl = cops[0].plan
    out[name] = dict(chain_us=tc, grouped_us=tg, grouped_launches=len(gops), tiles=sum(o.plan.total_tiles for o in cops),
                     sim_us=sum(o.plan.sim_units for o in cops) * 0.44, ideal_us=sum(o.plan.ideal_units for o in cops) * 0.44,
                     grid=pl.grid)
    print(name, json.dumps(out[name]), flush=True)
