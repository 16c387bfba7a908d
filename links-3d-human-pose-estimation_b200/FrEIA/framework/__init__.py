"""``FrEIA.framework.SequenceINN`` on the fused sm_100a flow kernels."""
import torch
import torch.nn as nn

from links_b200 import _cabi
from links_b200.flowpack import FlowPacked

__all__ = ["SequenceINN"]


_TRAINABLE = ("subnet.0.weight", "subnet.0.bias", "subnet.2.weight", "subnet.2.bias", "global_scale", "global_offset")


class _FlowFn(torch.autograd.Function):
    """(z, log_jac_det) = inn(x).  `params`: the flow's trainable parameters in (block, _TRAINABLE) order when any of them
    requires grad (autograd then asks for their gradients: links_flow_vjp_train + the parameter-gradient GEMMs of
    links_b200.flowtrain), empty for a frozen flow (input gradient only: links_flow_vjp)."""

    @staticmethod
    def forward(ctx, inn, x, rev, *params):
        fp = inn._packed()
        xc = x.detach().contiguous().float()
        out, ld = fp.apply(xc, rev=rev)
        ctx.inn, ctx.rev, ctx.n_params = inn, rev, len(params)
        ctx.save_for_backward(xc, out)
        return out, ld

    @staticmethod
    def backward(ctx, gz, gld):
        xc, out = ctx.saved_tensors
        inn = ctx.inn
        if ctx.rev:
            raise NotImplementedError("gradients through the reverse pass are not needed by any LInKs step "
                                      "(sampling runs under torch.no_grad, train_leg_torso_lifter.py:133)")
        gz = torch.zeros_like(xc) if gz is None else gz.contiguous().float()
        gld = None if gld is None else gld.contiguous().float()
        need_x = ctx.needs_input_grad[1]
        if ctx.n_params and any(ctx.needs_input_grad[3:]):
            eng = inn._grad_engine(xc.shape[0])
            dx = torch.empty_like(xc) if need_x else None
            eng.vjp(xc, gz, gld, dx)
            grads = [eng.Gd[k][n].clone() if ctx.needs_input_grad[3 + k * len(_TRAINABLE) + i] else None
                     for k in range(eng.nb) for i, n in enumerate(_TRAINABLE)]
            return (None, dx, None) + tuple(grads)
        if not need_x:
            return (None, None, None) + (None,) * ctx.n_params
        fp = inn._packed()
        dx = torch.empty_like(xc)
        _cabi.check(fp.lib.links_flow_vjp(fp.packed.data_ptr(), fp.C, fp.n_blocks, xc.data_ptr(), xc.shape[0],
                                          gz.data_ptr(), gld.data_ptr() if gld is not None else None, dx.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "links_flow_vjp")
        return (None, dx, None) + (None,) * ctx.n_params


class SequenceINN(nn.Module):
    """``inn = SequenceINN(C); inn.append(AllInOneBlock, subnet_constructor=..., permute_soft=True)`` ...
    ``z, log_jac_det = inn(x)``, ``x, log_jac_det = inn(z, rev=True)``.  The forward call is differentiable like FrEIA's:
    gradients flow to the input and, for parameters with requires_grad=True, to the flow's parameters (what
    train_full_pose_norm_flow.py:75-98 relies on).  ``inn.input_grad_only = True`` skips the parameter gradients (the
    lifter trainers never read them: no optimiser owns their flows, train_leg_torso_lifter.py:109-121); the fused
    training step ``links_b200.flowtrain.FlowTrainStep`` remains the fast way to train a flow."""

    input_grad_only = False

    def __init__(self, *dims, force_tuple_output=False):
        super().__init__()
        self.shapes = [tuple(dims)]
        self.conditions = []
        self.module_list = nn.ModuleList()
        self.force_tuple_output = force_tuple_output
        self._pk = None
        self._pk_ver = None
        self._geng = {}            # rows -> links_b200.flowtrain.FlowTrainStep serving parameter gradients

    def append(self, module_class, cond=None, cond_shape=None, **kwargs):
        if cond is not None:
            raise NotImplementedError("conditioning is not used by LInKs")
        module = module_class([self.shapes[-1]], **kwargs)
        self.module_list.append(module)
        self.shapes.append(self.shapes[-1])
        self._pk = None
        self._geng = {}

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

    def _trainable(self):
        named = dict(self.named_parameters())
        return [named["module_list.%d.%s" % (k, n)] for k in range(len(self.module_list)) for n in _TRAINABLE]

    def _grad_engine(self, rows):
        """Workspaces + plans of the parameter-gradient path for `rows` rows, holding the CURRENT parameter values."""
        from links_b200.flowtrain import FlowTrainStep
        ps = self._trainable()
        ver = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)
        eng = self._geng.get(rows)
        if eng is None:
            if len(self._geng) >= 4:                 # a handful of batch sizes at most (workspaces are 4 KB per row and block)
                self._geng.pop(next(iter(self._geng)))
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            eng = FlowTrainStep(self.shapes[0][0], sd, max(rows // 2, 1), n_blocks=len(self.module_list),
                                device=ps[0].device, external_rows=True, rows=rows)
            self._geng[rows] = eng
        elif eng._ver != ver:
            with torch.no_grad():
                for k in range(eng.nb):
                    for i, n in enumerate(_TRAINABLE):
                        eng.P[k][n].copy_(ps[k * len(_TRAINABLE) + i])
            eng._refresh_shadows()
        eng._ver = ver
        return eng

    def forward(self, x_or_z, c=None, rev=False, jac=True, force_tuple_output=False):
        if not x_or_z.is_cuda:
            raise _cabi.LinksError("links_b200 SequenceINN runs on a B200 only (no CPU fallback)")
        params = ()
        if torch.is_grad_enabled() and not rev and not self.input_grad_only:
            ps = self._trainable()
            if any(p.requires_grad for p in ps):
                params = tuple(ps)
        out, ld = _FlowFn.apply(self, x_or_z, bool(rev), *params)
        return ((out,), ld) if (self.force_tuple_output or force_tuple_output) else (out, ld)
