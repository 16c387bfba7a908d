"""Device-side packed parameters of a FrEIA-style flow (SequenceINN of AllInOneBlocks) and kernel launches.

State-dict keys follow FrEIA: ``module_list.{k}.{global_scale,global_offset,w_perm,w_perm_inv,
subnet.0.weight,subnet.0.bias,subnet.2.weight,subnet.2.bias}`` (reference call sites
train_leg_torso_lifter.py:352-367)."""
import ctypes as C

import torch

from . import _cabi
from ._cabi import check

_NAMES = ("subnet.0.weight", "subnet.0.bias", "subnet.2.weight", "subnet.2.bias", "global_scale", "global_offset",
          "w_perm", "w_perm_inv")


class FlowPacked:
    def __init__(self, C_dim, params, n_blocks=8, device="cuda"):
        self.C, self.n_blocks = C_dim, n_blocks
        self.lib = _cabi.lib()
        self.device = torch.device(device)
        n = self.lib.links_flow_packed_floats(C_dim, n_blocks)
        if n == 0:
            raise ValueError("unsupported flow width %d / %d blocks" % (C_dim, n_blocks))
        self.packed = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.repack(params)

    def repack(self, params):
        """(Re)build the packed buffer from FrEIA-layout tensors (device or host)."""
        self._src = {n: [params["module_list.%d.%s" % (k, n)].detach().to(self.device, torch.float32).contiguous()
                         for k in range(self.n_blocks)] for n in _NAMES}
        tables = [(C.c_void_p * self.n_blocks)(*[t.data_ptr() for t in self._src[n]]) for n in _NAMES]
        check(self.lib.links_flow_pack(self.C, self.n_blocks, *tables, self.packed.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "links_flow_pack")

    def apply(self, x, rev=False, out=None, ld=None):
        M = x.shape[0]
        out = torch.empty_like(x) if out is None else out
        ld = torch.empty(M, dtype=torch.float32, device=x.device) if ld is None else ld
        check(self.lib.links_flow_apply(self.packed.data_ptr(), self.C, self.n_blocks, x.data_ptr(), M, 1 if rev else 0,
                                        out.data_ptr(), ld.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "links_flow_apply")
        return out, ld

    def stash_for(self, rows):
        """Activation stash for nll_fwdbwd on `rows` rows (kept per row count; see links_flow_nll_fwdbwd)."""
        st = getattr(self, "_stash", None)
        if st is None or st[0] != rows:
            n = self.lib.links_flow_stash_floats(self.C, self.n_blocks, rows)
            self._stash = st = (rows, torch.empty(n, dtype=torch.float32, device=self.device))
        return st[1]

    def nll_fwdbwd(self, x, scale, nll_sum, dx, stash=True):
        sp = self.stash_for(x.shape[0]).data_ptr() if stash else None
        check(self.lib.links_flow_nll_fwdbwd(self.packed.data_ptr(), self.C, self.n_blocks, x.data_ptr(), x.shape[0],
                                             scale, nll_sum.data_ptr(), dx.data_ptr() if dx is not None else None, sp,
                                             torch.cuda.current_stream().cuda_stream), "links_flow_nll_fwdbwd")

    def sample(self, x, noise, out):
        check(self.lib.links_flow_sample(self.packed.data_ptr(), self.n_blocks, x.data_ptr(), noise.data_ptr(),
                                         x.shape[0], out.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "links_flow_sample")
        return out
