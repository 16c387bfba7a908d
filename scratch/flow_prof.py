import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "links-3d-human-pose-estimation_b200")]
import torch
from links_b200.flowpack import FlowPacked
from oracle import flow as OF
Cdim, M = 34, 16384
params = OF.init_flow_params(Cdim, 50 + Cdim, perturb=0.3)
fp = FlowPacked(Cdim, params)
x = (torch.randn(M, Cdim) * 0.2).cuda()
nll_sum = torch.zeros(1, device="cuda"); dx = torch.zeros(M, Cdim, device="cuda")
for _ in range(3):
    fp.apply(x)
    fp.nll_fwdbwd(x, 1.0 / M, nll_sum, dx)
torch.cuda.synchronize()
print("ok")
