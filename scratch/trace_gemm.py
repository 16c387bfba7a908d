"""Debug: per-CTA phase timestamps of the grouped GEMM (scratch/tracelib build with -DLINKS_GEMM_TRACE)."""
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
        if k == "sign_out": P.ld_sign = v.stride(0)
        elif k == "bits": P.ld_bits = v.stride(0)
        elif k == "out_f32": P.ld_f32 = v.stride(0)
        elif k != "bias": setattr(P, "ld_" + k, v.stride(0))
    return P

M, N, K = 2048, 1024, 1024
def mk(nprob, epi):
    probs, keep = [], []
    for i in range(nprob):
        A = (torch.randn(M, K, device="cuda") * 0.3).bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.randn(N, device="cuda")
        kw = dict(out=out, bias=bias)
        if epi:
            resid = (torch.randn(M, N, device="cuda") * 0.3).bfloat16(); sign = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
            kw.update(add0=resid, sign_out=sign, flags=_cabi.EPI_LEAKY_PRE | _cabi.EPI_LEAKY_POST); keep += [resid, sign]
        keep += [A, W, out, bias]
        probs.append(prob(A, W, M, N, K, **kw))
    return (_cabi.GemmProblem * nprob)(*probs), keep

st = torch.cuda.current_stream().cuda_stream
for nprob, epi, dbg in ((2, False, 0), (2, False, 1 << 30), (2, False, 1 << 29), (2, False, 3 << 29), (4, True, 0)):
    arr, keep = mk(nprob, epi)
    for P in arr: P.flags |= dbg
    print('dbg flags', hex(dbg))
    for _ in range(5): L.links_gemm_grouped(arr, nprob, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.links_gemm_grouped(arr, nprob, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    fl = 2.0 * M * N * K * nprob
    tr = np.zeros(148 * 16, np.uint64)
    L.links_debug_gemm_trace(tr.ctypes.data)
    tr = tr.reshape(148, 16).astype(np.int64)
    t0 = tr[:, 0].min()
    rel = (tr - t0) / 1000.0
    print("nprob %d epi %d: %.1f us/launch  %.0f TFLOP/s" % (nprob, epi, us, fl / us / 1e6))
    names = ["start", "setup_done", "first_full", "t0_accfull", "t0_epi_end", "-", "t1_accfull", "t1_epi_end", "-", "t2_accfull", "t2_epi_end", "", "", "", "stores_done", "exit"]
    for cta in (0, 100, 147 if nprob > 2 else 127):
        print("   cta %3d:" % cta, " ".join("%s=%.1f" % (names[i], rel[cta, i]) for i in (0, 1, 2, 3, 4, 6, 7, 9, 10, 14, 15) if tr[cta, i] >= t0))

# ---- small-N launches (heads, upscale dgrad, upscale fwd K=64)
def timeit(arr, n, reps=20):
    for _ in range(5): L.links_gemm_grouped(arr, n, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): L.links_gemm_grouped(arr, n, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / reps
keep2 = []
def heads(nets):
    ps = []
    for Nh in nets:
        A = (torch.randn(M, 1024, device="cuda") * 0.3).bfloat16(); W = (torch.randn(Nh, 1024, device="cuda") * 0.03).bfloat16()
        out = torch.zeros(M, 32, device="cuda"); bias = torch.randn(64, device="cuda")[:Nh]
        keep2.extend([A, W, out, bias])
        ps.append(prob(A, W, M, Nh, 1024, out_f32=out, bias=bias))
    return (_cabi.GemmProblem * len(ps))(*ps), len(ps)
arr, n = heads([7, 1, 10, 1]); print("heads fwd (7,1,10,1): %.1f us" % timeit(arr, n))
arr, n = heads([11, 11]); print("heads fwd (11,11): %.1f us" % timeit(arr, n))
ps = []
for kin in (14, 20):
    Gm = (torch.randn(M, 1024, device="cuda") * 0.3).bfloat16(); W = torch.zeros(1024, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(M, 32, device="cuda"); keep2.extend([Gm, W, out])
    P = prob(Gm, W, M, kin, 1024, out_f32=out, flags=_cabi.GEMM_B_MN); ps.append(P)
arr = (_cabi.GemmProblem * 2)(*ps); print("upscale dgrad: %.1f us" % timeit(arr, 2))
ps = []
for kin in (14, 20):
    X = torch.zeros(M, 64, device="cuda", dtype=torch.bfloat16); W = (torch.randn(1024, 64, device="cuda") * 0.03).bfloat16()
    out = torch.zeros(M, 1024, device="cuda", dtype=torch.bfloat16); bias = torch.randn(1024, device="cuda"); keep2.extend([X, W, out, bias])
    ps.append(prob(X, W, M, 1024, 64, out=out, bias=bias))
arr = (_cabi.GemmProblem * 2)(*ps); print("upscale fwd K=64: %.1f us" % timeit(arr, 2))
# wgrad 8 problems 1024x1024 over 4096 rows
ps = []
for i in range(8):
    Gm = (torch.randn(4096, 1024, device="cuda") * 0.1).bfloat16(); X = (torch.randn(4096, 1024, device="cuda") * 0.3).bfloat16()
    out = torch.zeros(1024, 1024, device="cuda"); keep2.extend([Gm, X, out])
    ps.append(prob(Gm, X, 1024, 1024, 4096, out_f32=out, flags=_cabi.GEMM_A_MN | _cabi.GEMM_B_MN))
arr = (_cabi.GemmProblem * 8)(*ps); t = timeit(arr, 8); print("wgrad 8x(1024x1024x4096): %.1f us  %.0f TFLOP/s" % (t, 8 * 2 * 1024 * 1024 * 4096 / t / 1e6))
