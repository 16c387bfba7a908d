from links_b200 import _cabi
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

