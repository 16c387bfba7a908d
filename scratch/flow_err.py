import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "links-3d-human-pose-estimation_b200")]
import torch, numpy as np
from links_b200.flowpack import FlowPacked
from oracle import flow as OF
for Cdim, M in [(14, 200), (20, 512), (22, 256), (34, 2048)]:
    params = OF.init_flow_params(Cdim, 50 + Cdim, perturb=0.3)
    fp = FlowPacked(Cdim, params)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(M, Cdim, generator=g) * 0.2
    p64 = {k: v.double() for k, v in params.items()}
    xg = x.double().clone().requires_grad_(True)
    zz, ll = OF.inn_forward(xg, p64)
    nll = OF.nll(zz, ll); (nll.sum() / M).backward()
    gref = xg.grad.numpy()
    xg32 = x.clone().requires_grad_(True)
    z32, l32 = OF.inn_forward(xg32, params); (OF.nll(z32, l32).sum() / M).backward()
    xd = x.cuda()
    for simt in (0, 1):
        fp.lib.links_flow_set_simt_only(simt)
        nll_sum = torch.zeros(1, device="cuda"); dx = torch.zeros(M, Cdim, device="cuda")
        fp.nll_fwdbwd(xd, 1.0 / M, nll_sum, dx)
        z, ld = fp.apply(xd)
        torch.cuda.synchronize()
        e = np.abs(dx.cpu().numpy() - gref).max() / np.abs(gref).max()
        ez = np.abs(z.cpu().numpy() - zz.detach().numpy()).max()
        el = np.abs(ld.cpu().numpy() - ll.detach().numpy()).max()
        print(Cdim, M, "simt" if simt else "tc  ", "grad err/max %.2e" % e, "z abs %.2e" % ez, "ld abs %.2e" % el, "nll rel %.2e" % abs(nll_sum.item() / nll.sum().item() - 1))
    e32 = np.abs(xg32.grad.numpy() - gref).max() / np.abs(gref).max()
    print(Cdim, M, "cpu32", "grad err/max %.2e" % e32, "z abs %.2e" % np.abs(z32.detach().numpy() - zz.detach().numpy()).max())
    fp.lib.links_flow_set_simt_only(0)
# timing
import time
for Cdim, M in [(14, 2048), (20, 2048), (22, 2048), (34, 1024), (34, 16384)]:
    params = OF.init_flow_params(Cdim, 50 + Cdim, perturb=0.3)
    fp = FlowPacked(Cdim, params)
    x = (torch.randn(M, Cdim) * 0.2).cuda()
    nll_sum = torch.zeros(1, device="cuda"); dx = torch.zeros(M, Cdim, device="cuda")
    for simt in (0, 1):
        fp.lib.links_flow_set_simt_only(simt)
        for name, fn in (("nll_fwdbwd", lambda: fp.nll_fwdbwd(x, 1.0 / M, nll_sum, dx)), ("fwd", lambda: fp.apply(x))):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            print(Cdim, M, "simt" if simt else "tc  ", name, "%.1f us" % (e0.elapsed_time(e1) * 100))
    fp.lib.links_flow_set_simt_only(0)
