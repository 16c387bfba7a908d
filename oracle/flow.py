"""Oracle (test infrastructure): FrEIA SequenceINN + AllInOneBlock restatement.

PARITY UNPINNED: FrEIA (github.com/VLL-HD/FrEIA; ``FrEIA.framework.SequenceINN``,
``FrEIA.modules.AllInOneBlock``) is a third-party dependency the reference neither
vendors nor pins, and it is not installed here.  This file restates its published
algorithm for the only configuration the reference uses
(``append(Fm.AllInOneBlock, subnet_constructor=subnet_fc, permute_soft=True)``;
call sites train_leg_torso_lifter.py:352-363, train_left_right_lifter.py:521-527,
train_full_pose_norm_flow.py:47-49): affine_clamping=2.0 with tanh clamp,
global_affine_init=1.0, global_affine_type='SOFTPLUS' (0.1*softplus_{beta=0.5}),
gin_block=False, no conditioning, 1-D inputs.

Parameters: ``dict[str, Tensor]`` with FrEIA's state-dict keys
``module_list.{k}.{global_scale,global_offset,w_perm,w_perm_inv,subnet.0.weight,
subnet.0.bias,subnet.2.weight,subnet.2.bias}``.
"""
import math

import numpy as np
import torch

CLAMP = 2.0
N_BLOCKS = 8
HIDDEN = 1024


def splits(C):
    return C - C // 2, C // 2  # (c1, c2): FrEIA split_len1 = ceil, split_len2 = floor


def _softplus_half(x):
    # nn.Softplus(beta=0.5, threshold=20)
    return torch.nn.functional.softplus(x, beta=0.5)


def subnet(x1, p, pre):
    """subnet_fc: Linear(c1,1024) -> ReLU -> Linear(1024, 2*c2), helpers.py:291-293."""
    h = torch.relu(torch.nn.functional.linear(x1, p[pre + "subnet.0.weight"], p[pre + "subnet.0.bias"]))
    return torch.nn.functional.linear(h, p[pre + "subnet.2.weight"], p[pre + "subnet.2.bias"])


def block_forward(x, p, k, rev=False):
    """One AllInOneBlock.  Returns (out, log_jac_det[M])."""
    pre = "module_list.%d." % k
    C = x.shape[1]
    c1, c2 = splits(C)
    g = 0.1 * _softplus_half(p[pre + "global_scale"])          # [1, C]
    perm_log_jac = torch.log(g).sum()
    if rev:
        x = (torch.nn.functional.linear(x, p[pre + "w_perm_inv"]) - p[pre + "global_offset"]) / g
    x1, x2 = x[:, :c1], x[:, c1:]
    a = subnet(x1, p, pre) * 0.1
    s = CLAMP * torch.tanh(a[:, :c2])
    t = a[:, c2:]
    if not rev:
        y2 = x2 * torch.exp(s) + t
        j = s.sum(dim=1)
    else:
        y2 = (x2 - t) * torch.exp(-s)
        j = -s.sum(dim=1)
    out = torch.cat((x1, y2), dim=1)
    if not rev:
        out = torch.nn.functional.linear(out * g + p[pre + "global_offset"], p[pre + "w_perm"])
        j = j + perm_log_jac
    else:
        j = j - perm_log_jac
    return out, j


def inn_forward(x, p, rev=False, n_blocks=N_BLOCKS):
    """SequenceINN.__call__(x, rev) -> (out, log_jac_det)."""
    ld = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    order = range(n_blocks - 1, -1, -1) if rev else range(n_blocks)
    for k in order:
        x, j = block_forward(x, p, k, rev=rev)
        ld = ld + j
    return x, ld


def nll(z, ld):
    """0.5*sum z^2 - log_jac_det (train_full_pose_norm_flow.py:77)."""
    return 0.5 * torch.sum(z ** 2, 1) - ld


def init_flow_params(C, seed, dtype=torch.float32, n_blocks=N_BLOCKS, perturb=0.0):
    """FrEIA-style init.  w_perm from scipy special_ortho_group (seeded here; FrEIA uses
    the NumPy global RNG).  ``perturb`` > 0 moves global_scale/global_offset off their
    init so tests exercise them (a trained checkpoint would have arbitrary values)."""
    from scipy.stats import special_ortho_group
    gen = torch.Generator().manual_seed(seed)
    c1, c2 = splits(C)
    p = {}
    gs0 = 2.0 * math.log(math.exp(0.5 * 10.0 * 1.0) - 1.0)
    for k in range(n_blocks):
        pre = "module_list.%d." % k
        gs = torch.full((1, C), gs0, dtype=torch.float64)
        go = torch.zeros((1, C), dtype=torch.float64)
        if perturb:
            gs = gs + perturb * torch.randn(1, C, generator=gen, dtype=torch.float64)
            go = go + 0.1 * perturb * torch.randn(1, C, generator=gen, dtype=torch.float64)
        w = special_ortho_group.rvs(C, random_state=np.random.RandomState(seed * 131 + k))
        w = torch.from_numpy(np.asarray(w, dtype=np.float64))
        p[pre + "global_scale"] = gs.to(dtype)
        p[pre + "global_offset"] = go.to(dtype)
        p[pre + "w_perm"] = w.to(dtype).contiguous()
        p[pre + "w_perm_inv"] = w.t().to(dtype).contiguous()
        b0 = 1.0 / math.sqrt(c1)
        p[pre + "subnet.0.weight"] = ((torch.rand(HIDDEN, c1, generator=gen, dtype=torch.float64) * 2 - 1) * b0).to(dtype)
        p[pre + "subnet.0.bias"] = ((torch.rand(HIDDEN, generator=gen, dtype=torch.float64) * 2 - 1) * b0).to(dtype)
        b2 = 1.0 / math.sqrt(HIDDEN)
        p[pre + "subnet.2.weight"] = ((torch.rand(2 * c2, HIDDEN, generator=gen, dtype=torch.float64) * 2 - 1) * b2).to(dtype)
        p[pre + "subnet.2.bias"] = ((torch.rand(2 * c2, generator=gen, dtype=torch.float64) * 2 - 1) * b2).to(dtype)
    return p
