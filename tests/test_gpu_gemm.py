"""GPU: the tcgen05 grouped GEMM through the C ABI vs a torch fp32 reference on the same bf16-rounded operands
(a floating-point kernel: torch reference allowed by the tier contract)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from links_b200 import _cabi
    return _cabi, _cabi.lib()


def _prob(cabi, A, B, M, N, K, **kw):
    P = cabi.GemmProblem()
    P.A, P.B, P.M, P.N, P.K, P.lda, P.ldb = A.data_ptr(), B.data_ptr(), M, N, K, A.stride(0), B.stride(0)
    P.flags = kw.pop("flags", 0)
    for k, v in kw.items():
        if k in ("ld_f32", "outT_col0"):
            setattr(P, k, v)
        elif v is not None:
            setattr(P, k, v.data_ptr())
            if k == "outT":
                P.ld_outT = v.stride(0)
            elif k == "out_f32":
                P.ld_f32 = v.stride(0)
            elif k == "sign_out":
                P.ld_sign = v.stride(0)
            elif k == "bits":
                P.ld_bits = v.stride(0)
            elif k != "bias":
                setattr(P, "ld_" + k, v.stride(0))
    return P


def _run(cabi, L, probs):
    arr = (cabi.GemmProblem * len(probs))(*probs)
    rc = L.links_gemm_grouped(arr, len(probs), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, rc


def leaky(x):
    return torch.where(x > 0, x, 0.01 * x)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (2048, 1024, 1024), (200, 1024, 1024), (77, 7, 1024),
                                   (1024, 14, 300), (130, 1024, 64)])
def test_plain_gemm_f32_out(M, N, K):
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    lda, ldb = (K + 7) // 8 * 8, (K + 7) // 8 * 8
    A = (torch.randn(M, lda, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, ldb, device="cuda", generator=g) * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    P = _prob(cabi, A, B, M, N, K, bias=bias, out_f32=out)
    _run(cabi, L, [P])
    ref = A[:, :K].float() @ B[:, :K].float().t() + bias
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-3)
    # accumulate mode
    P = _prob(cabi, A, B, M, N, K, out_f32=out, flags=cabi.EPI_ACCUM_F32)
    _run(cabi, L, [P])
    torch.testing.assert_close(out, 2 * ref - bias, rtol=2e-4, atol=4e-3)


def test_forward_epilogues_and_transpose():
    cabi, L = _lib()
    M, N, K = 300, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.3).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = (torch.randn(M, N, device="cuda", generator=g) * 0.3).bfloat16()
    acc = A.float() @ W.float().t() + bias
    ldT = 312
    # l1-style: leaky(acc + b)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    outT = torch.zeros(N, ldT, device="cuda", dtype=torch.bfloat16)
    _run(cabi, L, [_prob(cabi, A, W, M, N, K, bias=bias, flags=cabi.EPI_LEAKY_PRE, out=out, outT=outT, outT_col0=8)])
    ref = leaky(acc)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-3)
    assert torch.equal(outT[:, 8:8 + M], out.t())
    assert torch.all(outT[:, :8] == 0)
    # an unaligned column offset is rejected (TMA stores need 16-byte aligned starts)
    P = _prob(cabi, A, W, M, N, K, out=out, outT=outT, outT_col0=5)
    assert L.links_gemm_grouped((cabi.GemmProblem * 1)(P), 1, None) == -2
    # l2-style: leaky(leaky(acc + b) + resid) with the sign mask of (acc + b)
    sign = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
    _run(cabi, L, [_prob(cabi, A, W, M, N, K, bias=bias, flags=cabi.EPI_LEAKY_PRE | cabi.EPI_LEAKY_POST, add0=resid,
                         sign_out=sign, out=out)])
    ref = leaky(leaky(acc) + resid.float())
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-3)
    bits = ((sign.unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, N).bool()
    expect = ~(acc > 0)
    # bits may differ only where acc is within accumulation-order noise of zero
    assert ((bits != expect) & (acc.abs() > 1e-3)).sum().item() == 0
    assert (bits != expect).float().mean().item() < 1e-3


def test_backward_epilogue_masks():
    cabi, L = _lib()
    M, N, K = 256, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(2)
    G = (torch.randn(M, K, device="cuda", generator=g) * 0.1).bfloat16()
    WT = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    skip = (torch.randn(M, N, device="cuda", generator=g) * 0.1).bfloat16()
    extra = (torch.randn(M, N, device="cuda", generator=g) * 0.1).bfloat16()
    y = (torch.randn(M, N, device="cuda", generator=g)).bfloat16()
    bits_bool = torch.rand(M, N, device="cuda", generator=g) < 0.5
    w = (bits_bool.reshape(M, N // 32, 32).long() << torch.arange(32, device="cuda")).sum(-1)
    bits = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()
    mid = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    outT = torch.zeros(N, M, device="cuda", dtype=torch.bfloat16)
    _run(cabi, L, [_prob(cabi, G, WT, M, N, K, add0=skip, add1=extra, ymask=y, mid=mid, bits=bits, out=out, outT=outT)])
    v = G.float() @ WT.float().t() + skip.float() + extra.float()
    v = v * torch.where(y.float() > 0, 1.0, 0.01)
    torch.testing.assert_close(mid.float(), v, rtol=1e-2, atol=1e-3)
    v = v * torch.where(bits_bool, 0.01, 1.0)
    torch.testing.assert_close(out.float(), v, rtol=1e-2, atol=1e-3)
    assert torch.equal(outT, out.t())


def test_grouped_heterogeneous_problems():
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(2048, 1024, 1024), (512, 1024, 64), (2048, 11, 1024), (100, 1, 1024), (1024, 1024, 4096)]
    probs, refs, outs = [], [], []
    keep = []
    for (M, N, K) in shapes:
        A = (torch.randn(M, K, device="cuda", generator=g) * 0.2).bfloat16()
        B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        out = torch.zeros(M, 32 if N < 32 else N, device="cuda")
        keep += [A, B]
        probs.append(_prob(cabi, A, B, M, N, K, out_f32=out))
        outs.append(out[:, :N])
        refs.append(A.float() @ B.float().t())
    _run(cabi, L, probs)
    for o, r in zip(outs, refs):
        torch.testing.assert_close(o, r, rtol=2e-4, atol=3e-3)


def test_argument_errors():
    cabi, L = _lib()
    A = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    P = _prob(cabi, A, A, 128, 128, 64)
    P.lda = 60   # not a multiple of 8
    arr = (cabi.GemmProblem * 1)(P)
    assert L.links_gemm_grouped(arr, 1, None) == -2
    assert L.links_gemm_grouped(arr, 0, None) == -1
    with pytest.raises(ValueError):
        cabi.check(-2, "x")
