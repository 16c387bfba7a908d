"""ncu target: each geometry kernel (LT maps, N = 4 M rows, packed heads) twice, the fused eval kernel twice."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "links-3d-human-pose-estimation_b200"))
from links_b200 import _cabi, maps  # noqa: E402

L = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream
N = 4 * 1024 * 1024
f32 = dict(dtype=torch.float32, device="cuda")
kind = sys.argv[1] if len(sys.argv) > 1 else "lt"
m = maps.geom_maps(kind)
nj = (7, 10) if kind == "lt" else (11, 11)
u = torch.randn(N, 34, **f32) * 0.1
pack1, pack2 = torch.randn(N, 32, **f32) * 0.1, torch.randn(N, 32, **f32) * 0.1
heads = [pack1[:, 0:], pack1[:, nj[0]:]]
angs = [pack1[:, 24:], pack1[:, 25:]]
heads2 = [pack2[:, 0:], pack2[:, nj[0]:]]
eps, uy = torch.randn(N, **f32), torch.rand(N, **f32)
stats = torch.zeros(2, **f32)
qp = [torch.zeros(N, 2 * nj[s], **f32) for s in range(2)]
common = [u.data_ptr(), heads[0].data_ptr(), heads[1].data_ptr(), angs[0].data_ptr(), angs[1].data_ptr(), eps.data_ptr(),
          uy.data_ptr(), stats.data_ptr()]
_cabi.check(L.links_elev_stats(angs[0].data_ptr(), angs[1].data_ptr(), N, stats.data_ptr(), st), "stats")
sums = torch.zeros(4, **f32)
g2 = [torch.zeros(N, 64, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
dfl = [torch.randn(N, 2 * nj[s], **f32) for s in range(2)]
dli = [torch.randn(N, 32, **f32) for _ in range(2)]
g1 = [torch.zeros(N, 64, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
dgam, da, red = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(2, **f32)
for _ in range(2):
    L.links_geom_forward(C.byref(m), *common, N, qp[0].data_ptr(), qp[1].data_ptr(), None, None, st)
    L.links_geom_loss(C.byref(m), *common, heads2[0].data_ptr(), heads2[1].data_ptr(), N, sums.data_ptr(),
                      g2[0].data_ptr(), g2[1].data_ptr(), None, None, 0, 0, st)
    L.links_geom_backward(C.byref(m), *common, heads2[0].data_ptr(), heads2[1].data_ptr(), dfl[0].data_ptr(),
                          dfl[1].data_ptr(), dli[0].data_ptr(), dli[1].data_ptr(), N, g1[0].data_ptr(),
                          g1[1].data_ptr(), None, None, 0, 0, dgam.data_ptr(), da.data_ptr(), red.data_ptr(), st)
torch.cuda.synchronize()
M = 4 * 1024 * 1024
gt = torch.randn(M, 51, **f32) * 300
p2d = torch.randn(M, 34, **f32) * 0.1
doff = torch.randn(M, 32, **f32) * 0.1
s3 = torch.zeros(3, dtype=torch.float64, device="cuda")
for _ in range(2):
    L.links_eval_lift_score(p2d.data_ptr(), doff.data_ptr(), 32, gt.data_ptr(), M, 10.0, s3.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
