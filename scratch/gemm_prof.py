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
M, N, K = 2048, 1024, 1024
probs, keep = [], []
for i in range(2):
    A = (torch.randn(M, K, device="cuda") * 0.3).bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.randn(N, device="cuda")
    keep += [A, W, out, bias]
    probs.append(prob(A, W, M, N, K, out=out, bias=bias))
arr = (_cabi.GemmProblem * 2)(*probs)
st = torch.cuda.current_stream().cuda_stream
for _ in range(6): L.links_gemm_grouped(arr, 2, st)
torch.cuda.synchronize()
print("ok")
