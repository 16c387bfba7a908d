"""Module-level drop-in (utils.models_def.Leg_Lifter as an nn.Module under torch autograd + torch.optim.Adam) at 2 048 rows:
forward, forward + backward, forward + backward + optimiser step.  BASELINE.md quotes the reference on CPU: 126 ms forward,
450 ms forward + backward."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
import torch
from utils.models_def import Leg_Lifter

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
net = Leg_Lifter(use_batchnorm=False, num_joints=7, use_dropout=False).cuda()
opt = torch.optim.Adam(net.parameters(), lr=2e-4, weight_decay=1e-5)
x = torch.randn(M, 14, device="cuda") * 0.1


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd():
    with torch.no_grad():
        return net(x)


def fwd_bwd():
    opt.zero_grad(set_to_none=True)
    d, a = net(x)
    (d.square().mean() + a.square().mean()).backward()


def train():
    fwd_bwd()
    opt.step()


flops_fwd = 2.0 * M * (14 * 1024 + 14 * 1024 * 1024 + 1024 * 8)
out = {"rows": M}
for name, fn, mult in (("forward", fwd, 1.0), ("forward_backward", fwd_bwd, 3.0), ("train_step_torch_adam", train, 3.0)):
    ms = timed(fn)
    out[name] = {"ms": ms, "tflops": flops_fwd * mult / (ms * 1e-3) / 1e12}
print(json.dumps(out))
