"""Random-init parameters with the reference's distributions and state-dict key names.

nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias;
reference utils/models_def.py builds every layer from nn.Linear) and FrEIA AllInOneBlock init
(global_scale = 2*log(exp(5)-1), global_offset = 0, w_perm = scipy special_ortho_group, subnet = subnet_fc
with default nn.Linear init; reference utils/helpers.py:291-293).  Deterministic in `seed`.
"""
import math

import numpy as np
import torch

LIFTER_BLOCKS = ("res_common", "res_pose1", "res_pose2", "res_pose3", "res_angle1", "res_angle2", "res_angle3")
PREDICTOR_BLOCKS = ("res_common", "res_pose1", "res_pose2", "res_pose3")


def _linear(gen, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w.float(), b.float()


def _mlp(in_dim, heads, blocks, seed, width=1024):
    gen = torch.Generator().manual_seed(seed)
    p = {}
    p["upscale.weight"], p["upscale.bias"] = _linear(gen, width, in_dim)
    for blk in blocks:
        for l in ("l1", "l2"):
            p["%s.%s.weight" % (blk, l)], p["%s.%s.bias" % (blk, l)] = _linear(gen, width, width)
    for name, od in heads.items():
        p[name + ".weight"], p[name + ".bias"] = _linear(gen, od, width)
    return p


def init_lifter_params(num_joints, seed):
    return _mlp(2 * num_joints, {"downscale": num_joints, "angles": 1}, LIFTER_BLOCKS, seed)


def init_predictor_params(num_joints_in, out_dim, seed):
    return _mlp(3 * num_joints_in, {"downscale": out_dim}, PREDICTOR_BLOCKS, seed)


def init_flow_params(C, seed, n_blocks=8, perturb=0.3, hidden=1024):
    from scipy.stats import special_ortho_group
    gen = torch.Generator().manual_seed(seed)
    c1, c2 = C - C // 2, C // 2
    gs0 = 2.0 * math.log(math.exp(5.0) - 1.0)
    p = {}
    for k in range(n_blocks):
        pre = "module_list.%d." % k
        gs = torch.full((1, C), gs0, dtype=torch.float64)
        go = torch.zeros((1, C), dtype=torch.float64)
        if perturb:   # a trained flow has arbitrary global affine values; keep the synthetic one off its init
            gs = gs + perturb * torch.randn(1, C, generator=gen, dtype=torch.float64)
            go = go + 0.1 * perturb * torch.randn(1, C, generator=gen, dtype=torch.float64)
        w = torch.from_numpy(np.asarray(special_ortho_group.rvs(C, random_state=np.random.RandomState(seed * 131 + k)),
                                        dtype=np.float64))
        p[pre + "global_scale"], p[pre + "global_offset"] = gs.float(), go.float()
        p[pre + "w_perm"], p[pre + "w_perm_inv"] = w.float().contiguous(), w.t().float().contiguous()
        b0, b2 = 1.0 / math.sqrt(c1), 1.0 / math.sqrt(hidden)
        p[pre + "subnet.0.weight"] = ((torch.rand(hidden, c1, generator=gen, dtype=torch.float64) * 2 - 1) * b0).float()
        p[pre + "subnet.0.bias"] = ((torch.rand(hidden, generator=gen, dtype=torch.float64) * 2 - 1) * b0).float()
        p[pre + "subnet.2.weight"] = ((torch.rand(2 * c2, hidden, generator=gen, dtype=torch.float64) * 2 - 1) * b2).float()
        p[pre + "subnet.2.bias"] = ((torch.rand(2 * c2, generator=gen, dtype=torch.float64) * 2 - 1) * b2).float()
    return p
