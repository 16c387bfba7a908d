"""``FrEIA.framework.SequenceINN`` on the fused sm_100a flow kernels."""
import torch
import torch.nn as nn

from links_b200 import _cabi
from links_b200.flowpack import FlowPacked

__all__ = ["SequenceINN"]


class _FlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inn, x, rev):
        fp = inn._packed()
        xc = x.detach().contiguous().float()
        out, ld = fp.apply(xc, rev=rev)
        ctx.inn, ctx.rev = inn, rev
        ctx.save_for_backward(xc, out)
        return out, ld

    @staticmethod
    def backward(ctx, gz, gld):
        xc, out = ctx.saved_tensors
        fp = ctx.inn._packed()
        if ctx.rev:
            raise NotImplementedError("gradients through the reverse pass are not needed by any LInKs step "
                                      "(sampling runs under torch.no_grad, train_leg_torso_lifter.py:133)")
        gz = torch.zeros_like(xc) if gz is None else gz.contiguous().float()
        gld = None if gld is None else gld.contiguous().float()
        dx = torch.empty_like(xc)
        _cabi.check(fp.lib.links_flow_vjp(fp.packed.data_ptr(), fp.C, fp.n_blocks, xc.data_ptr(), xc.shape[0],
                                          gz.data_ptr(), gld.data_ptr() if gld is not None else None, dx.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "links_flow_vjp")
        return None, dx, None


class SequenceINN(nn.Module):
    """``inn = SequenceINN(C); inn.append(AllInOneBlock, subnet_constructor=..., permute_soft=True)`` ...
    ``z, log_jac_det = inn(x)``, ``x, log_jac_det = inn(z, rev=True)``.  Gradients flow to the *input*
    (frozen flows, as in the lifter trainers); training the flow's own parameters goes through
    ``links_b200.flowtrain.FlowTrainStep`` (train_full_pose_norm_flow.py drop-in)."""

    _warned = False

    def __init__(self, *dims, force_tuple_output=False):
        super().__init__()
        self.shapes = [tuple(dims)]
        self.conditions = []
        self.module_list = nn.ModuleList()
        self.force_tuple_output = force_tuple_output
        self._pk = None
        self._pk_ver = None

    def append(self, module_class, cond=None, cond_shape=None, **kwargs):
        if cond is not None:
            raise NotImplementedError("conditioning is not used by LInKs")
        module = module_class([self.shapes[-1]], **kwargs)
        self.module_list.append(module)
        self.shapes.append(self.shapes[-1])
        self._pk = None

    def __len__(self):
        return len(self.module_list)

    def __getitem__(self, i):
        return self.module_list[i]

    def _packed(self):
        ps = list(self.parameters())
        ver = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)
        if self._pk is None or self._pk_ver != ver:
            sd = {k: v for k, v in self.state_dict().items()}
            dev = ps[0].device
            if dev.type != "cuda":
                raise _cabi.LinksError("links_b200 SequenceINN runs on a B200 only (no CPU fallback): call .cuda()")
            C = self.shapes[0][0]
            if self._pk is None:
                self._pk = FlowPacked(C, sd, n_blocks=len(self.module_list), device=dev)
            else:
                self._pk.repack(sd)
            self._pk_ver = ver
        return self._pk

    def forward(self, x_or_z, c=None, rev=False, jac=True, force_tuple_output=False):
        if not x_or_z.is_cuda:
            raise _cabi.LinksError("links_b200 SequenceINN runs on a B200 only (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # Every lifter trainer of the reference calls its flows with requires_grad=True parameters that NO optimiser
            # owns (train_leg_torso_lifter.py:109-121,203-216): autograd computes their gradients and nobody reads them.
            # The shim therefore propagates the INPUT gradient only and leaves .grad of the flow parameters None.  Code
            # that trains the flow through this call (train_full_pose_norm_flow.py:75) must use
            # links_b200.flowtrain.FlowTrainStep (what the drop-in train_full_pose_norm_flow.py does); strict=True turns
            # this case into an error instead of a warning.
            if getattr(self, "strict_param_grads", False):
                raise NotImplementedError("parameter gradients of the flow are provided by links_b200.flowtrain."
                                          "FlowTrainStep; freeze the flow or wrap the call in torch.no_grad()")
            if not SequenceINN._warned:
                SequenceINN._warned = True
                import warnings
                warnings.warn("links_b200 FrEIA shim: flow parameters require grad, but this call only propagates "
                              "gradients to its INPUT (flow parameters keep .grad = None).  Train flows with "
                              "links_b200.flowtrain.FlowTrainStep.", stacklevel=2)
        out, ld = _FlowFn.apply(self, x_or_z, bool(rev))
        return ((out,), ld) if (self.force_tuple_output or force_tuple_output) else (out, ld)
