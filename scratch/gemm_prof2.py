import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "links-3d-human-pose-estimation_b200")]
import torch
from links_b200 import _cabi
L = _cabi.lib()
def prob(A, B, M, N, K, **kw):
    P = _cabi.GemmProblem()
    P.A, P.B, P.M, P.N, P.K, P.lda, P.ldb = A.data_ptr(), B.data_ptr(), M, N, K, A.stride(0), B.stride(0)
    P.flags = kw.pop("flags", 0)
    for k, v in kw.items():
        setattr(P, k, v.data_ptr())
        if k == "sign_out": P.ld_sign = v.stride(0)
        elif k != "bias": setattr(P, "ld_" + k, v.stride(0))
    return P
M, N, K = 16384, 1024, 1024
probs, keep = [], []
for i in range(4):
    A = (torch.randn(M, K, device="cuda") * 0.3).bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.randn(N, device="cuda")
    resid = (torch.randn(M, N, device="cuda") * 0.3).bfloat16(); sign = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
    keep += [A, W, out, bias, resid, sign]
    probs.append(prob(A, W, M, N, K, out=out, bias=bias, add0=resid, sign_out=sign, flags=_cabi.EPI_LEAKY_PRE | _cabi.EPI_LEAKY_POST))
arr = (_cabi.GemmProblem * 4)(*probs)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): L.links_gemm_grouped(arr, 4, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): L.links_gemm_grouped(arr, 4, st)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / 5
print("4 x (16384x1024x1024) mode3: %.1f us  %.0f TFLOP/s" % (us, 4 * 2.0 * M * N * K / us / 1e6))
