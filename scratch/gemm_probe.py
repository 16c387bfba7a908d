"""Probe GEMM epilogue variants: python scratch/gemm_probe.py <case>"""
import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'links-3d-human-pose-estimation_b200'); sys.path.insert(0, 'tests')
from links_b200 import _cabi
from test_gpu_gemm import _prob, _run, leaky
cabi, L = _cabi, _cabi.lib()
case = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(1)


def mk(M, N, K):
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.3).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    return A, W


if case.startswith("t"):
    # correctness of transposed / row-major stores: t,<M>,<col0>,<ldT>
    _, M, col0, ldT = case.split(",")
    M, col0, ldT = int(M), int(col0), int(ldT)
    N, K = 1024, 256
    A, W = mk(M, N, K)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    outT = torch.full((N, ldT), 7.0, device="cuda", dtype=torch.bfloat16)
    _run(cabi, L, [_prob(cabi, A, W, M, N, K, out=out, outT=outT, outT_col0=col0)])
    ref = (A.float() @ W.float().t())
    print(case, "out err", (out.float() - ref).abs().max().item(), "T equal", torch.equal(outT[:, col0:col0 + M], out.t()),
          "untouched left", bool((outT[:, :col0] == 7).all()), "untouched right", bool((outT[:, col0 + M:] == 7).all()))
else:
    # timing: which outputs cost what (4 problems of 2048x1024x1024)
    M, N, K = 2048, 1024, 1024
    probs, keep = [], []
    for i in range(4):
        A, W = mk(M, N, K)
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        outT = torch.zeros(N, 4096, device="cuda", dtype=torch.bfloat16)
        f32 = torch.zeros(M, N, device="cuda")
        add = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        keep += [A, W, out, outT, f32, add]
        kw = {}
        if "o" in case: kw["out"] = out
        if "T" in case: kw["outT"] = outT
        if "f" in case: kw["out_f32"] = f32
        if "a" in case: kw["add0"] = add
        if "m" in case: kw["mid"] = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); keep.append(kw["mid"])
        probs.append(_prob(cabi, A, W, M, N, K, **kw))
    arr = (cabi.GemmProblem * 4)(*probs)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        L.links_gemm_grouped(arr, 4, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        L.links_gemm_grouped(arr, 4, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(case, "%.1f us  %.0f TFLOP/s" % (us, 4 * 2 * M * N * K / us / 1e6))
