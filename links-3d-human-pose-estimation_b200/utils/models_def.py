"""Drop-in for reference ``utils/models_def.py``: same class names, constructor arguments, forward signatures and
``state_dict`` keys (including the never-used LayerNorm ``bn1/bn2`` of every res_block and the unused
``res_common`` of the occlusion predictors), with forward AND backward running on the sm_100a tcgen05 GEMM engine
(``links_b200.mlp.MlpSet``) through a ``torch.autograd.Function``.

There is no CPU fallback: inputs must be CUDA tensors on a B200.  ``use_batchnorm`` / ``use_dropout`` are accepted
for signature compatibility but must stay False -- every reference call site passes False
(e.g. train_leg_torso_lifter.py:76-77, train_occlusion_models.py:90-97).
"""
import torch
import torch.nn as nn

from links_b200 import _cabi
from links_b200.mlp import MlpSet

__all__ = ["res_block", "PoseDiscriminator", "DepthAngleEstimator", "Leg_Lifter", "Torso_Lifter", "Left_Right_Lifter",
           "Occluded_Limb_Predictor", "Occluded_Legs_Predictor", "Occluded_Torso_Predictor",
           "Occluded_Left_Right_Predictor"]


class res_block(nn.Module):
    """Parameter container of reference models_def.py:10-39 (l1, bn1, d1, l2, bn2, d2).  Its arithmetic,
    LeakyReLU(l2(LeakyReLU(l1(x)))) + x, is executed fused inside the owning network's GEMM epilogues."""

    def __init__(self, num_neurons: int = 1024, use_batchnorm: bool = False, use_dropout: bool = False, dropout=0.5):
        super().__init__()
        self.use_batchnorm = use_batchnorm
        self.use_dropout = use_dropout
        self.l1 = nn.Linear(num_neurons, num_neurons)
        self.bn1 = nn.LayerNorm(num_neurons)
        self.d1 = nn.Dropout(float(dropout))
        self.l2 = nn.Linear(num_neurons, num_neurons)
        self.bn2 = nn.LayerNorm(num_neurons)
        self.d2 = nn.Dropout(float(dropout))

    def forward(self, x):
        raise _cabi.LinksError("res_block is evaluated fused inside its parent network on the B200 engine; "
                               "call the parent module instead")


class _EngineFn(torch.autograd.Function):
    """y_heads = net(x): forward / dgrad / wgrad through the grouped tcgen05 GEMM plans of a 1-network MlpSet."""

    @staticmethod
    def forward(ctx, module, train, x, *params):
        # grad mode is always off inside Function.forward: the caller decides whether activations are kept
        eng = module._engine(x.shape[0], train=train)
        module._sync_params(eng)
        if train:
            # this autograd node owns the workspace's activations until its backward has run (the reference calls
            # every lifter twice, every predictor three times per step before one backward:
            # train_leg_torso_lifter.py:150-151,227-228; train_occlusion_models.py:196-300)
            eng._gen += 1
            eng._busy = True
        ctx.gen = eng._gen
        M, K = x.shape
        lib = eng.lib
        st = torch.cuda.current_stream().cuda_stream
        xc = x.detach().contiguous().float()
        idx = module._identity_index(K, x.device)
        _cabi.check(lib.links_pack_rows(xc.data_ptr(), K, M, idx.data_ptr(), K, 1, eng.x0[0][0].data_ptr(), None, 0, 0, st),
                    "links_pack_rows")
        eng.run(eng.forward_ops(0))
        outs = []
        for head, width in module._heads:
            outs.append(eng.head_out[0][0][head][:, :width].clone())
        ctx.module, ctx.eng, ctx.M, ctx.K = module, eng, M, K
        ctx.need_x = x.requires_grad
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        module, eng, M, K = ctx.module, ctx.eng, ctx.M, ctx.K
        if not eng.train:
            raise _cabi.LinksError("forward ran under no_grad; no activations were kept for backward")
        if ctx.gen != eng._gen:
            raise _cabi.LinksError("the activations of this forward call were recycled: more than %d forward calls of one "
                                   "module (same batch size) were alive without a backward; raise "
                                   "utils.models_def.MAX_LIVE_FORWARDS" % MAX_LIVE_FORWARDS)
        for (head, width), g in zip(module._heads, gouts):
            G = eng.G[0][0][head]
            G.zero_()
            if g is not None:
                G[:, :width] = g.to(torch.bfloat16)
        eng.run(eng.backward_ops(0, need_input_grad=ctx.need_x, wgrad=True))     # dgrad chain + weight / bias gradients
        gx = eng.din[0][0][:, :K].clone() if ctx.need_x else None
        grads = []
        for name in module._param_order:
            layer, kind = name.rsplit(".", 1)
            L = eng.nets[0].layers[layer]
            grads.append((L.gW if kind == "weight" else L.gb).clone())
        eng._busy = False
        return (None, None, gx) + tuple(grads)


MAX_LIVE_FORWARDS = 8     # activation workspaces kept per (module, batch size): forward calls alive before a backward


class _EngineModule(nn.Module):
    """Shared machinery: lazily built engines (a small pool of activation workspaces per batch size that share ONE set
    of parameter / gradient / shadow buffers), parameter sync by version counter."""
    _kind = "lifter"

    def _post_init(self, in_dim, heads, use_batchnorm, use_dropout):
        if use_batchnorm or use_dropout:
            raise NotImplementedError("the B200 path implements the configuration every reference script uses: "
                                      "use_batchnorm=False, use_dropout=False")
        self._in_dim = in_dim
        self._heads = heads                      # [(name, width)]
        self._engines = {}
        self._versions = {}
        self._idx = {}

    def _used_layers(self):
        trunk, branches = {"lifter": (["res_common"], ["res_pose1", "res_pose2", "res_pose3", "res_angle1", "res_angle2",
                                                       "res_angle3"]),
                           "predictor": ([], ["res_pose1", "res_pose2", "res_pose3"])}[self._kind]
        names = ["upscale"]
        for blk in trunk + branches:
            names += [blk + ".l1", blk + ".l2"]
        return names + [h for h, _ in self._heads]

    @property
    def _param_order(self):
        return [n + s for n in self._used_layers() for s in (".weight", ".bias")]

    def _identity_index(self, K, device):
        if K not in self._idx:
            self._idx[K] = torch.arange(K, dtype=torch.int32, device=device)
        return self._idx[K]

    def _engine(self, rows, train):
        """A workspace whose activations no pending backward needs.  Workspaces of one (rows, train) key alias the
        parameter buffers of the first one; at most MAX_LIVE_FORWARDS exist, after that the one whose forward is oldest
        is recycled (its generation counter moves on, so a late backward of the evicted call raises instead of silently
        using the wrong activations)."""
        key = (rows, bool(train))
        pool = self._engines.setdefault(key, [])
        for eng in pool:
            if not eng._busy:
                return eng
        if len(pool) < MAX_LIVE_FORWARDS:
            eng = MlpSet(self._kind, [self._in_dim], [dict(self._heads)], rows, n_passes=1,
                         device=next(self.parameters()).device, train=bool(train), share_from=pool[0] if pool else None)
            eng._gen, eng._busy, eng._key = 0, False, key
            pool.append(eng)
            if len(pool) == 1:
                self._versions[key] = None
            return eng
        eng = min(pool, key=lambda e: e._gen)
        return eng

    def _sync_params(self, eng):
        sd = dict(self.named_parameters())
        ver = tuple(sd[n]._version for n in self._param_order) + tuple(sd[n].data_ptr() for n in self._param_order)
        key = eng._key
        if self._versions.get(key) != ver:
            eng.load_state_dicts([{n: sd[n].detach() for n in self._param_order}])     # buffers shared by the pool
            self._versions[key] = ver

    def _run(self, x):
        if not x.is_cuda:
            raise _cabi.LinksError("links_b200 modules run on a B200 only (no CPU fallback); move the module and "
                                   "its input to cuda")
        sd = dict(self.named_parameters())
        params = [sd[n] for n in self._param_order]
        shape = x.shape
        train = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        out = _EngineFn.apply(self, train, x.reshape(-1, shape[-1]), *params)
        return out


def _make_blocks(mod, names, use_batchnorm, use_dropout, dropout):
    for n in names:
        setattr(mod, n, res_block(use_batchnorm=use_batchnorm, num_neurons=1024, use_dropout=use_dropout, dropout=dropout))


class _Lifter(_EngineModule):
    """Leg_/Torso_/Left_Right_Lifter (models_def.py:111-239): forward(x[M, 2J]) -> (depths[M, J], angle[M, 1])."""
    _kind = "lifter"

    def __init__(self, use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.5):
        super().__init__()
        self.upscale = nn.Linear(2 * num_joints, 1024)
        _make_blocks(self, ["res_common", "res_pose1", "res_pose2", "res_pose3", "res_angle1", "res_angle2", "res_angle3"],
                     use_batchnorm, use_dropout, use_dropout)      # the reference passes dropout=use_dropout (:116)
        self.downscale = nn.Linear(1024, num_joints)
        self.angles = nn.Linear(1024, 1)
        self._post_init(2 * num_joints, [("downscale", num_joints), ("angles", 1)], use_batchnorm, use_dropout)

    def forward(self, x):
        xd, xa = self._run(x)
        return xd, xa


class Leg_Lifter(_Lifter):
    def __init__(self, use_batchnorm=False, num_joints=7, use_dropout=False, d_rate=0.5):
        super().__init__(use_batchnorm, num_joints, use_dropout, d_rate)


class Torso_Lifter(_Lifter):
    def __init__(self, use_batchnorm=False, num_joints=10, use_dropout=False, d_rate=0.5):
        super().__init__(use_batchnorm, num_joints, use_dropout, d_rate)


class Left_Right_Lifter(_Lifter):
    def __init__(self, use_batchnorm=False, num_joints=11, use_dropout=False, d_rate=0.5):
        super().__init__(use_batchnorm, num_joints, use_dropout, d_rate)


class DepthAngleEstimator(_Lifter):
    """ElePose leftover (models_def.py:65-107), same topology as the lifters with `dropout` instead of `d_rate`."""

    def __init__(self, use_batchnorm=False, num_joints=16, use_dropout=False, dropout=0.5):
        super().__init__(use_batchnorm, num_joints, use_dropout, dropout)


class _Predictor(_EngineModule):
    """Occluded_*_Predictor (models_def.py:243-327): forward(x[M, 3*num_joints]) -> [M, out]; res_common unused."""
    _kind = "predictor"
    _out = 9

    def __init__(self, use_batchnorm=False, num_joints=10):
        super().__init__()
        self.upscale = nn.Linear(3 * num_joints, 1024)
        _make_blocks(self, ["res_common", "res_pose1", "res_pose2", "res_pose3"], use_batchnorm, False, 0.5)
        self.downscale = nn.Linear(1024, self._out)
        self._post_init(3 * num_joints, [("downscale", self._out)], use_batchnorm, False)

    def forward(self, x):
        return self._run(x)[0]


class Occluded_Limb_Predictor(_Predictor):
    _out = 3 * 3


class Occluded_Legs_Predictor(_Predictor):
    _out = 3 * 6


class Occluded_Torso_Predictor(_Predictor):
    _out = 3 * 10


class Occluded_Left_Right_Predictor(_Predictor):
    _out = 3 * 6


class PoseDiscriminator(nn.Module):
    """ElePose leftover (models_def.py:42-63): upscale -> LeakyReLU(res_common) -> downscale.  Unused by every
    script; kept as a parameter-compatible container (state-dict keys identical), not accelerated."""

    def __init__(self, use_batchnorm=False, num_joints=16, use_dropout=False, dropout=0.5):
        super().__init__()
        self.upscale = nn.Linear(2 * num_joints, 1024)
        _make_blocks(self, ["res_common", "res_pose1", "res_pose2"], use_batchnorm, use_dropout, dropout)
        self.downscale = nn.Linear(1024, 1)

    def forward(self, x):
        raise _cabi.LinksError("PoseDiscriminator is outside the accelerated hot path (no reference script uses it)")
