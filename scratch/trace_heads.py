import os, sys, ctypes as C
os.environ["LINKS_B200_LIB"] = os.path.join(os.getcwd(), "scratch/tracelib/liblinks_b200.so")
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "links-3d-human-pose-estimation_b200")]
import numpy as np, torch
from links_b200 import _cabi
L = _cabi.lib()
L.links_debug_gemm_trace.restype = C.c_int
L.links_debug_gemm_trace.argtypes = [C.c_void_p]
def prob(A, B, M, N, K, **kw):
    P = _cabi.GemmProblem()
    P.A, P.B, P.M, P.N, P.K, P.lda, P.ldb = A.data_ptr(), B.data_ptr(), M, N, K, A.stride(0), B.stride(0)
    P.flags = kw.pop("flags", 0)
    for k, v in kw.items():
        setattr(P, k, v.data_ptr())
        if k == "out_f32": P.ld_f32 = v.stride(0)
        elif k != "bias": setattr(P, "ld_" + k, v.stride(0))
    return P
M = 2048
st = torch.cuda.current_stream().cuda_stream
keep = []
def run(name, ps):
    arr = (_cabi.GemmProblem * len(ps))(*ps)
    for _ in range(5): L.links_gemm_grouped(arr, len(ps), st)
    torch.cuda.synchronize()
    tr = np.zeros(148 * 16, np.uint64); L.links_debug_gemm_trace(tr.ctypes.data)
    tr = tr.reshape(148, 16).astype(np.int64); t0 = tr[0, 0]
    names = {0: "start", 1: "setup", 2: "first_full", 3: "accfull", 4: "epi_end", 15: "exit"}
    print(name, " ".join("%s=%.1f" % (names[i], (tr[0, i] - t0) / 1000.0) for i in (0, 1, 2, 3, 4, 15)))
ps = []
for Nh in (7, 10):
    A = (torch.randn(M, 1024, device="cuda") * 0.3).bfloat16(); W = (torch.randn(Nh, 1024, device="cuda") * 0.03).bfloat16()
    out = torch.zeros(M, 32, device="cuda"); bias = torch.randn(64, device="cuda")[:Nh]; keep += [A, W, out, bias]
    ps.append(prob(A, W, M, Nh, 1024, out_f32=out, bias=bias))
run("heads", ps)
ps = []
for kin in (14, 20):
    Gm = (torch.randn(M, 1024, device="cuda") * 0.3).bfloat16(); W = torch.zeros(1024, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(M, 32, device="cuda"); keep += [Gm, W, out]
    ps.append(prob(Gm, W, M, kin, 1024, out_f32=out, flags=_cabi.GEMM_B_MN))
run("updgrad", ps)
