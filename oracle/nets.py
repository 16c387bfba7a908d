"""Oracle (test infrastructure): residual-MLP lifters / occlusion predictors.

Functional restatement of reference ``utils/models_def.py``.  Parameters are plain
``dict[str, Tensor]`` keyed exactly like the reference ``state_dict()`` so the same
tensors load into the reference modules, this oracle and the CUDA product.
"""
import math

import torch

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, models_def.py:28,34

LIFTER_BLOCKS = ("res_common", "res_pose1", "res_pose2", "res_pose3",
                 "res_angle1", "res_angle2", "res_angle3")
PREDICTOR_BLOCKS = ("res_common", "res_pose1", "res_pose2", "res_pose3")


def leaky(x):
    return torch.nn.functional.leaky_relu(x, LEAKY_SLOPE)


def linear(x, p, name):
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


def res_block(x, p, name):
    """models_def.py:23-39 with use_batchnorm=use_dropout=False (the only mode the scripts use)."""
    y = leaky(linear(x, p, name + ".l1"))
    y = leaky(linear(y, p, name + ".l2"))
    return y + x


def lifter_forward(x, p, pose_only=False):
    """Leg_/Torso_/Left_Right_Lifter.forward, models_def.py:133-152 (identical in all three).

    ``pose_only`` skips the angle branch (its output is discarded by every caller that
    passes ``_`` -- pass 2, validation, eval, occlusion)."""
    h = linear(x, p, "upscale")
    hc = leaky(res_block(h, p, "res_common"))
    xd = hc
    for k in (1, 2, 3):
        xd = leaky(res_block(xd, p, "res_pose%d" % k))
    xd = linear(xd, p, "downscale")
    if pose_only:
        return xd, None
    xa = hc
    for k in (1, 2, 3):
        xa = leaky(res_block(xa, p, "res_angle%d" % k))
    xa = linear(xa, p, "angles")
    return xd, xa


def predictor_forward(x, p):
    """Occluded_*_Predictor.forward, models_def.py:253-263 (res_common unused)."""
    h = linear(x, p, "upscale")
    xd = h
    for k in (1, 2, 3):
        xd = leaky(res_block(xd, p, "res_pose%d" % k))
    return linear(xd, p, "downscale")


def _init_linear(gen, out_f, in_f, dtype):
    # nn.Linear.reset_parameters: kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for both
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w.to(dtype), b.to(dtype)


def init_mlp_params(in_dim, out_dims, blocks, seed, dtype=torch.float32, width=1024):
    """Random-init parameters with nn.Linear's distribution and the reference key names.

    out_dims: dict head name -> width, e.g. {"downscale": 7, "angles": 1}."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    p["upscale.weight"], p["upscale.bias"] = _init_linear(gen, width, in_dim, dtype)
    for blk in blocks:
        for l in ("l1", "l2"):
            p["%s.%s.weight" % (blk, l)], p["%s.%s.bias" % (blk, l)] = _init_linear(gen, width, width, dtype)
    for name, od in out_dims.items():
        p[name + ".weight"], p[name + ".bias"] = _init_linear(gen, od, width, dtype)
    return p


def init_lifter_params(num_joints, seed, dtype=torch.float32):
    return init_mlp_params(2 * num_joints, {"downscale": num_joints, "angles": 1}, LIFTER_BLOCKS, seed, dtype)


def init_predictor_params(num_joints_in, out_dim, seed, dtype=torch.float32):
    return init_mlp_params(3 * num_joints_in, {"downscale": out_dim}, PREDICTOR_BLOCKS, seed, dtype)
