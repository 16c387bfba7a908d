"""Oracle (test infrastructure): the lifter/predictor networks with the CUDA path's BF16 rounding points.

Same algorithm as ``oracle.nets`` (reference utils/models_def.py), but every GEMM operand and every stored
activation / activation-gradient is rounded to bfloat16 exactly where the sm_100a kernels round
(fp32 accumulate, fp32 bias / head outputs / weight gradients).  Gradient checks against this twin are tight;
the fp32 oracle stays the parity target for joints and losses (north star: 1e-3 relative).
"""
import torch

from . import nets as N32


def rb(x):
    return x.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    """bf16 storage of an activation: rounds in forward, passes gradients through."""

    @staticmethod
    def forward(ctx, x):
        return rb(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """bf16 storage of an activation gradient: identity in forward, rounds in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return rb(g)


class _QLinear(torch.autograd.Function):
    """y = bf16(x) bf16(W)^T + b with fp32 accumulation; backward GEMMs take bf16 operands too."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return rb(x) @ rb(w).t() + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gq = rb(g)
        return gq @ rb(w), gq.t() @ rb(x), gq.sum(0)


rf, rg = _RoundFwd.apply, _RoundGrad.apply


def qlinear(x, p, name):
    return _QLinear.apply(x, p[name + ".weight"], p[name + ".bias"])


def res_block(x, p, name):
    """x is a stored (bf16-valued) activation.  Returns the stored output of LeakyReLU(res_block(x))."""
    a1 = rf(N32.leaky(rg(qlinear(x, p, name + ".l1"))))
    z2 = rg(qlinear(a1, p, name + ".l2"))
    t = N32.leaky(z2) + rg(x)            # skip path: its gradient is stored in bf16 ("dt") before being re-added
    return rf(N32.leaky(t))


def lifter_forward(x, p, pose_only=False):
    h0 = rf(rg(qlinear(x, p, "upscale")))
    hc = res_block(h0, p, "res_common")
    xd = rg(hc) if not pose_only else hc   # pass-1 merge: the pose branch's contribution is stored (bf16) first
    for k in (1, 2, 3):
        xd = res_block(xd, p, "res_pose%d" % k)
    xd = qlinear(xd, p, "downscale")
    if pose_only:
        return xd, None
    xa = hc
    for k in (1, 2, 3):
        xa = res_block(xa, p, "res_angle%d" % k)
    return xd, qlinear(xa, p, "angles")


def predictor_forward(x, p):
    xd = rf(rg(qlinear(x, p, "upscale")))
    for k in (1, 2, 3):
        xd = res_block(xd, p, "res_pose%d" % k)
    return qlinear(xd, p, "downscale")
