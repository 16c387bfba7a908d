"""``FrEIA.modules.AllInOneBlock`` parameter container (arithmetic is fused inside SequenceINN's kernels)."""
import numpy as np
import torch
import torch.nn as nn

__all__ = ["AllInOneBlock"]


class AllInOneBlock(nn.Module):
    """Affine coupling + ActNorm-style global affine + (soft) permutation, FrEIA semantics:
    split x -> (x1[:c1], x2[c1:]), a = 0.1 * subnet(x1), s = clamp * tanh(a[:, :c2]), t = a[:, c2:],
    y2 = x2 * exp(s) + t, out = ((x1, y2) * g + offset) @ w_perm^T with g = 0.1 * softplus_{beta=0.5}(global_scale),
    log_jac_det = sum(s) + sum(log g).  Only the configuration LInKs uses is implemented on the GPU path:
    affine_clamping = 2.0, global_affine_type = 'SOFTPLUS', gin_block = False, no conditioning, 1-D inputs,
    subnet = Linear(c1, 1024) -> ReLU -> Linear(1024, 2 * c2)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, affine_clamping=2., gin_block=False,
                 global_affine_init=1., global_affine_type='SOFTPLUS', permute_soft=False,
                 learned_householder_permutation=0, reverse_permutation=False):
        super().__init__()
        if len(dims_c) or gin_block or learned_householder_permutation or reverse_permutation:
            raise NotImplementedError("links_b200 FrEIA shim: conditioning / GIN / householder permutations are not "
                                      "used by LInKs and not implemented")
        if global_affine_type != 'SOFTPLUS' or float(affine_clamping) != 2.0:
            raise NotImplementedError("links_b200 FrEIA shim implements global_affine_type='SOFTPLUS', "
                                      "affine_clamping=2.0 (the FrEIA defaults LInKs relies on)")
        if subnet_constructor is None:
            raise ValueError("Please supply a callable subnet_constructor function or object (see docstring)")
        channels = dims_in[0][0]
        if len(dims_in[0]) != 1:
            raise NotImplementedError("links_b200 FrEIA shim supports 1-D (fully connected) inputs only")
        self.in_channels = channels
        self.splits = [channels - channels // 2, channels // 2]
        self.clamp = affine_clamping
        global_scale = 2. * np.log(np.exp(0.5 * 10. * global_affine_init) - 1)
        self.global_scale = nn.Parameter(torch.ones(1, channels) * float(global_scale))
        self.global_offset = nn.Parameter(torch.zeros(1, channels))
        if permute_soft:
            from scipy.stats import special_ortho_group
            w = special_ortho_group.rvs(channels)         # NumPy global RNG, like FrEIA
        else:
            w = np.zeros((channels, channels))
            for i, j in enumerate(np.random.permutation(channels)):
                w[i, j] = 1.
        self.w_perm = nn.Parameter(torch.FloatTensor(w).view(channels, channels), requires_grad=False)
        self.w_perm_inv = nn.Parameter(torch.FloatTensor(w.T).view(channels, channels), requires_grad=False)
        self.subnet = subnet_constructor(self.splits[0], 2 * self.splits[1])
        ok = (isinstance(self.subnet, nn.Sequential) and len(self.subnet) == 3 and isinstance(self.subnet[0], nn.Linear)
              and isinstance(self.subnet[1], nn.ReLU) and isinstance(self.subnet[2], nn.Linear)
              and self.subnet[0].out_features == 1024 and self.subnet[2].in_features == 1024)
        if not ok:
            raise NotImplementedError("links_b200 FrEIA shim fuses subnet_fc = Linear(c1,1024)->ReLU->Linear(1024,2*c2) "
                                      "(reference utils/helpers.py:291-293); other subnets are not implemented")

    def output_dims(self, input_dims):
        return input_dims

    def forward(self, x, c=[], rev=False, jac=True):
        raise RuntimeError("AllInOneBlock is evaluated fused inside links_b200's SequenceINN")
