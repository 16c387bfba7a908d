"""Debug: per-tile phase timeline of a chain launch (trace build of the library: scratch/build_tracelib.sh)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("LINKS_B200_LIB", os.path.join(ROOT, "scratch/tracelib/liblinks_b200.so"))
sys.path[:0] = [ROOT, os.path.join(ROOT, "links-3d-human-pose-estimation_b200")]
import numpy as np, torch
from links_b200 import _cabi, init as INIT
from links_b200.mlp import MlpSet
L = _cabi.lib()
L.links_debug_chain_trace.restype = C.c_int
L.links_debug_chain_trace.argtypes = [C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
KT = 96
cases = [("fwd1", lambda: m._build_forward(1)), ("fwd0", lambda: m._build_forward(0)), ("bwd1", lambda: m._build_backward(1, True)),
         ("bwd0+wgrad", lambda: m._build_backward(0, False, None, True))]
for name, build in cases:
    ops = [op for op in m._chained(("t", name), build) if hasattr(op, "plan")]
    for _ in range(3):
        for op in ops: op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for op in ops: op()
    e1.record(); torch.cuda.synchronize()
    tr = np.zeros(74 * KT * 8, np.uint64)
    L.links_debug_chain_trace(tr.ctypes.data)
    tr = tr.reshape(74, KT, 8).astype(np.int64)
    pl = ops[0].plan
    ncl = pl.grid // 2
    t0 = tr[:ncl, 0, 1][tr[:ncl, 0, 1] > 0].min()
    print("== %s: %.1f us, %d tiles, %d clusters, sim %.1f us ideal %.1f us" % (name, e0.elapsed_time(e1) * 1e3, pl.total_tiles, ncl,
          pl.sim_units * 0.44, pl.ideal_units * 0.44))
    main, gap, epi, lag, depw, scout_lead = [], [], [], [], [], []
    ends = []
    for c in range(ncl):
        nt = int((tr[c, :, 1] >= t0).sum())
        for t in range(nt):
            a, b_, f, d = tr[c, t, 1], tr[c, t, 2], tr[c, t, 3], tr[c, t, 4]
            main.append((b_ - a) / 1e3)
            epi.append((d - f) / 1e3)
            lag.append((f - b_) / 1e3)
            if t > 0:
                gap.append((a - tr[c, t - 1, 2]) / 1e3)
                depw.append((tr[c, t, 0] - tr[c, t - 1, 2]) / 1e3)
        if nt:
            ends.append((tr[c, nt - 1, 4] - t0) / 1e3)
    f = lambda x: "mean %.2f p50 %.2f p90 %.2f max %.2f" % (np.mean(x), np.median(x), np.percentile(x, 90), np.max(x))
    print("   main loop (first operands -> last commit):", f(main))
    print("   MMA gap between tiles (last commit -> next first operands):", f(gap))
    print("   producer deps-cleared relative to previous tile's last commit:", f(depw))
    print("   accumulator-full lag after last commit:", f(lag))
    print("   epilogue (acc full -> tile done, first epilogue warp):", f(epi))
    print("   cluster finish times: min %.1f max %.1f" % (min(ends), max(ends)))
    # timeline of cluster 0
    c = 0
    nt = int((tr[c, :, 1] >= t0).sum())
    print("   cluster 0:", " | ".join("%.1f-%.1f e%.1f" % ((tr[c, t, 1] - t0) / 1e3, (tr[c, t, 2] - t0) / 1e3, (tr[c, t, 4] - t0) / 1e3) for t in range(min(nt, 14))))
