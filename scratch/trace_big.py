"""Debug: steady-state tile period vs epilogue duration of the grouped GEMM at large M (trace build)."""
import os, sys, ctypes as C
os.environ.setdefault("LINKS_B200_LIB", os.path.join(os.getcwd(), "scratch/tracelib/liblinks_b200.so"))
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "links-3d-human-pose-estimation_b200")]
import numpy as np, torch
from links_b200 import _cabi
L = _cabi.lib()
L.links_debug_gemm_trace.restype = C.c_int
L.links_debug_gemm_trace.argtypes = [C.c_void_p]
sys.path.insert(0, "scratch")
from trace_gemm_lib import prob  # noqa

def run(M, N, K, nprob, kind):
    probs, keep = [], []
    for i in range(nprob):
        A = (torch.randn(M, K, device="cuda") * 0.3).bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.randn(N, device="cuda")
        kw = dict(out=out, bias=bias)
        if kind == "l2":
            resid = (torch.randn(M, N, device="cuda") * 0.3).bfloat16(); sign = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
            kw.update(add0=resid, sign_out=sign, flags=_cabi.EPI_LEAKY_PRE | _cabi.EPI_LEAKY_POST); keep += [resid, sign]
        elif kind == "l1":
            kw.update(flags=_cabi.EPI_LEAKY_PRE)
        elif kind in ("d5", "d7"):
            del kw["bias"]
            ym = (torch.randn(M, N, device="cuda") * 0.3).bfloat16(); keep.append(ym)
            kw.update(ymask=ym)
            if kind == "d7":
                a0 = (torch.randn(M, N, device="cuda") * 0.3).bfloat16(); mid = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
                bits = torch.randint(0, 2 ** 31 - 1, (M, N // 32), device="cuda", dtype=torch.int32)
                keep += [a0, mid, bits]
                kw.update(add0=a0, mid=mid, bits=bits)
        keep += [A, W, out, bias]
        probs.append(prob(A, W, M, N, K, **kw))
    arr = (_cabi.GemmProblem * nprob)(*probs)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): L.links_gemm_grouped(arr, nprob, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.links_gemm_grouped(arr, nprob, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 10
    tr = np.zeros(148 * 16, np.uint64)
    L.links_debug_gemm_trace(tr.ctypes.data)
    tr = tr.reshape(148, 16).astype(np.int64)
    t0 = tr[:, 0].min()
    rel = (tr - t0) / 1000.0
    print("M %d nprob %d %s: %.1f us/launch  %.0f TFLOP/s" % (M, nprob, kind, us, 2.0 * M * N * K * nprob / us / 1e6))
    ok = tr[:, 10] >= t0
    if not ok.any():
        ok1 = tr[:, 4] >= t0
        print("   setup %.2f first_full %.2f t0_accfull %.2f epilogue(t0) %.2f exit %.2f" % (rel[:, 1].mean(), rel[::2, 2].mean(),
              rel[ok1, 3].mean(), (rel[ok1, 4] - rel[ok1, 3]).mean(), rel[:, 15].mean()))
        return
    per = (rel[ok, 9] - rel[ok, 6]); epi1 = rel[ok, 7] - rel[ok, 6]; epi0 = rel[ok, 4] - rel[ok, 3]
    print("   setup %.2f  first_full %.2f  t0_accfull %.2f | tile period (t2-t1 accfull) mean %.2f min %.2f max %.2f | epilogue(t0) %.2f epilogue(t1) %.2f (warp 2 only)" %
          (rel[:, 1].mean(), rel[::2, 2].mean(), rel[:, 3].mean(), per.mean(), per.min(), per.max(), epi0.mean(), epi1.mean()))

for M, nprob, kind in ((16384, 4, "l2"), (16384, 4, "d5"), (16384, 4, "d7"), (16384, 4, "plain"), (2048, 4, "l2"), (2048, 4, "d7"), (2048, 2, "l2")):
    run(M, 1024, 1024, nprob, kind)
