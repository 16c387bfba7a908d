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
        if k in ("ld_f32",):
            setattr(P, k, v)
        elif v is not None:
            setattr(P, k, v.data_ptr())
            if k == "out_f32":
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
    out = torch.full((M, N), 777.0, device="cuda")
    P = _prob(cabi, A, B, M, N, K, bias=bias, out_f32=out)
    _run(cabi, L, [P])
    ref = A[:, :K].float() @ B[:, :K].float().t() + bias
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-3)
    # accumulate mode
    P = _prob(cabi, A, B, M, N, K, out_f32=out, flags=cabi.EPI_ACCUM_F32)
    _run(cabi, L, [P])
    torch.testing.assert_close(out, 2 * ref - bias, rtol=2e-4, atol=4e-3)


def test_forward_epilogues():
    cabi, L = _lib()
    M, N, K = 300, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.3).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    resid = (torch.randn(M, N, device="cuda", generator=g) * 0.3).bfloat16()
    acc = A.float() @ W.float().t() + bias
    # l1-style: leaky(acc + b); the output buffer is wider than N (rows beyond M / columns beyond N stay untouched)
    big = torch.full((M + 5, N + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    out = big[:M, :N]
    _run(cabi, L, [_prob(cabi, A, W, M, N, K, bias=bias, flags=cabi.EPI_LEAKY_PRE, out=out)])
    ref = leaky(acc)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-3)
    assert torch.all(big[M:] == 7.0) and torch.all(big[:, N:] == 7.0)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
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
    _run(cabi, L, [_prob(cabi, G, WT, M, N, K, add0=skip, add1=extra, ymask=y, mid=mid, bits=bits, out=out)])
    v = G.float() @ WT.float().t() + skip.float() + extra.float()
    v = v * torch.where(y.float() > 0, 1.0, 0.01)
    torch.testing.assert_close(mid.float(), v, rtol=1e-2, atol=1e-3)
    v = v * torch.where(bits_bool, 0.01, 1.0)
    torch.testing.assert_close(out.float(), v, rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(256, 1024, 1024), (300, 1024, 11), (2048, 22, 1024), (77, 200, 130)])
def test_dgrad_mn_major_b(M, N, K):
    """dX[M, N] = G[M, K] . W[K, N] with W read in place as an MN-major B operand (no transposed shadow)."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    ldg, ldw = (K + 63) // 64 * 64, (N + 63) // 64 * 64
    G = torch.zeros(M, ldg, device="cuda", dtype=torch.bfloat16)
    G[:, :K] = (torch.randn(M, K, device="cuda", generator=g) * 0.3).bfloat16()
    W = torch.zeros(K, ldw, device="cuda", dtype=torch.bfloat16)
    W[:, :N] = (torch.randn(K, N, device="cuda", generator=g) * 0.1).bfloat16()
    out = torch.full((M, (N + 3) // 4 * 4), 777.0, device="cuda")
    P = _prob(cabi, G, W, M, N, K, flags=cabi.GEMM_B_MN, out_f32=out)
    P.ldb = ldw
    _run(cabi, L, [P])
    ref = G[:, :K].float() @ W[:, :N].float()
    torch.testing.assert_close(out[:, :N], ref, rtol=2e-4, atol=3e-3)


@pytest.mark.parametrize("rows,N,K", [(2048, 1024, 1024), (4096, 1024, 1024), (300, 11, 1024), (1000, 1024, 22), (130, 30, 42)])
def test_wgrad_mn_major_ab(rows, N, K):
    """dW[N, K] = G[rows, N]^T . X[rows, K]: both operands read in place as MN-major (contraction over rows)."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows + N + K)
    ldg, ldx = (N + 63) // 64 * 64, (K + 63) // 64 * 64
    G = torch.zeros(rows, ldg, device="cuda", dtype=torch.bfloat16)
    G[:, :N] = (torch.randn(rows, N, device="cuda", generator=g) * 0.1).bfloat16()
    X = torch.zeros(rows, ldx, device="cuda", dtype=torch.bfloat16)
    X[:, :K] = (torch.randn(rows, K, device="cuda", generator=g) * 0.3).bfloat16()
    out = torch.full((N, K), 777.0, device="cuda")
    P = _prob(cabi, G, X, N, K, rows, flags=cabi.GEMM_A_MN | cabi.GEMM_B_MN, out_f32=out)
    P.lda, P.ldb = ldg, ldx
    _run(cabi, L, [P])
    ref = G[:, :N].float().t() @ X[:, :K].float()
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=5e-3)


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


@pytest.mark.parametrize("M,passes", [(200, 1), (640, 1), (2048, 2)])
def test_chain_launch_equals_layer_by_layer(M, passes):
    """links_gemm_chain_run (one persistent kernel per pass, tile-level dependencies through completion counters) must
    reproduce the layer-by-layer grouped launches BIT FOR BIT: same tiles, same MMA order, same epilogues -- forward
    (heads, activations, sign masks), dgrad chain (all G buffers, input gradient) and the weight gradients.  Repeated
    launches of one plan exercise the counter reset by the last cluster to exit."""
    from links_b200.mlp import MlpSet
    from oracle import nets as ON
    nj = (7, 10)
    params = [ON.init_lifter_params(nj[0], 11), ON.init_lifter_params(nj[1], 12)]
    mlp = MlpSet("lifter", [14, 20], [{"downscale": 7, "angles": 1}, {"downscale": 10, "angles": 1}], M, n_passes=passes,
                 pass_branches=[["pose", "angle"], ["pose"]][:passes])
    mlp.load_state_dicts(params)
    g = torch.Generator().manual_seed(M)
    for p in range(passes):
        for s in range(2):
            mlp.x0[p][s].zero_()
            mlp.x0[p][s][:, :2 * nj[s]] = (torch.randn(M, 2 * nj[s], generator=g) * 0.2).bfloat16().cuda()

    def seed_head_grads():
        gg = torch.Generator().manual_seed(7)
        for p in range(passes):
            for s in range(2):
                for head, w in (("downscale", nj[s]), ("angles", 1)):
                    if head == "angles" and p == 1:
                        continue
                    G = mlp.G[p][s][head]
                    G.zero_()
                    G[:, :w] = (torch.randn(M, w, generator=gg) * 0.1).bfloat16().cuda()

    def snapshot():
        torch.cuda.synchronize()
        out = {}
        for p in range(passes):
            for s in range(2):
                for k, v in mlp.act[p][s].items():
                    out[("act", p, s, k)] = v.clone()
                for k, v in mlp.head_out[p][s].items():
                    out[("head", p, s, k)] = v.clone()
                for k, v in mlp.sign[p][s].items():
                    out[("sign", p, s, k)] = v.clone()
                for k, v in mlp.G[p][s].items():
                    out[("G", p, s, k)] = v.clone()
                out[("din", p, s)] = mlp.din[p][s].clone()
        for s in range(2):
            for n, L in mlp.nets[s].layers.items():
                out[("gW", s, n)] = L.gW.clone()
                out[("gb", s, n)] = L.gb.clone()
        return out

    def clear():
        for p in range(passes):
            for s in range(2):
                for v in mlp.act[p][s].values():
                    v.fill_(777.0)
                for v in mlp.head_out[p][s].values():
                    v.fill_(777.0)
                for k, v in mlp.G[p][s].items():
                    if k not in ("downscale", "angles"):
                        v.fill_(777.0)
                mlp.din[p][s].fill_(777.0)
                mlp.E[p][s].fill_(777.0)
                for v in mlp.dt[p][s].values():
                    v.fill_(777.0)
        mlp.grad.fill_(777.0)

    def run_all(chain):
        clear()
        seed_head_grads()
        last = passes - 1
        for p in range(passes):
            mlp.run(mlp._chained(("t_fwd", p), lambda p=p: mlp._build_forward(p)) if chain else mlp.forward_plan(p))
        order = list(range(passes - 1, -1, -1))          # the step runs the LAST pass's backward first
        for i, p in enumerate(order):
            wg = i == len(order) - 1 and p == 0
            ops = (mlp._chained(("t_bwd", p, wg), lambda p=p, wg=wg: mlp._build_backward(p, True, None, wg)) if chain
                   else mlp.backward_plan(p, True, None, wg))
            mlp.run(ops)
        return snapshot()

    ref = run_all(False)
    for rep in range(3):
        got = run_all(True)
        for k, v in ref.items():
            a, b = v, got[k]
            if k[0] == "gb":      # bias gradients: column sums with atomic partial sums (summation order varies run to run)
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (rep, k)
                continue
            assert torch.equal(a, b), (rep, k, (a.float() - b.float()).abs().max().item())
