"""Training step of a GLOW / AllInOne normalising flow on the B200 path (reference train_full_pose_norm_flow.py:67-98;
train_leg_torso_left_right_norm_flow.py:100-198 uses the same step on four part widths).

    z, ld = inn(x);                      dist_2d        = mean(0.5 |z|^2 - ld)                    (:75-78)
    s = inn^-1(add_noise(z)) (no grad), root joint zeroed                                        (:81-87)
    z2, ld2 = inn(s);                    dist_2d_sample = mean(0.5 |z2|^2 - ld2)                  (:89-91)
    loss = dist_2d + dist_2d_sample;  backward;  Adam                                            (:93-98)

Everything that depends on the rows runs in ONE fused tensor-core kernel launch over the stacked rows [x ; s]
(links_flow_nll_train): forward, NLL, reversible backward, gradients of the global affine, and the per-block operands
of the parameter-gradient GEMMs.  The weight gradients are then four grouped tcgen05 GEMM launches over the 8 blocks
(hidden recompute, dgrad with the ReLU mask, two wgrads), bias gradients one batched column-sum launch, Adam one launch
on the flat parameter buffer, and the packed operand images are rebuilt by links_flow_pack.
"""

import torch

from . import _cabi
from ._cabi import EPI_RELU_PRE, EPI_YMASK_ZERO, GEMM_A_MN, GEMM_B_MN, GemmProblem, check
from .flowpack import FlowPacked

HIDDEN = 1024
_NAMES = ("subnet.0.weight", "subnet.0.bias", "subnet.2.weight", "subnet.2.bias", "global_scale", "global_offset")


def _rup(x, m):
    return (x + m - 1) // m * m


class _GraphReplay:
    """step() behind a CUDA graph (the flow steps are 20-odd short launches: launch-bound when issued eagerly).
    run(): first call eager (lazy plans / workspaces), second call captures, later calls replay.  The inputs (x, noise)
    are static buffers refilled by the caller on the same stream; the learning rate is a device word (set_lr)."""
    graph = None
    _ran = 0

    def capture(self, warmup=1):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self.step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g
        return g

    def run(self, use_graph=True):
        if self.graph is not None:
            self.graph.replay()
        elif use_graph and self._ran >= 1 and self._graph_ok():
            self.capture(warmup=0).replay()
        else:
            self.step()
        self._ran += 1

    def _graph_ok(self):
        return True


class FlowTrainStep(_GraphReplay):
    def __init__(self, C_dim, params, batch, n_blocks=8, lr=2e-4, weight_decay=0.0, device="cuda", process_group=None,
                 external_rows=False, rows=None):
        """params: FrEIA-layout state dict; batch: rows of x per step (the kernel sees 2*batch rows).
        external_rows: the caller fills self.u ([2*batch, C] = [data rows ; sampled rows]) itself instead of the flow
        drawing samples from its own inverse (the part-flow trainer samples from a frozen full-pose flow).
        rows: row count of the workspaces when the object only serves vjp() (the FrEIA shim's autograd)."""
        self.external_rows = external_rows
        self.C, self.nb, self.B = C_dim, n_blocks, batch
        self.M = 2 * batch if rows is None else int(rows)
        self.c1, self.c2 = C_dim - C_dim // 2, C_dim // 2
        self.lr, self.wd = lr, weight_decay
        self.lib = _cabi.lib()
        self.device = dev = torch.device(device)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        c1, c2, nb, M = self.c1, self.c2, n_blocks, self.M
        # ---- flat trainable parameters (w_perm / w_perm_inv are fixed, FrEIA keeps them requires_grad=False)
        shapes = {"subnet.0.weight": (HIDDEN, c1), "subnet.0.bias": (HIDDEN,), "subnet.2.weight": (2 * c2, HIDDEN),
                  "subnet.2.bias": (2 * c2,), "global_scale": (1, C_dim), "global_offset": (1, C_dim)}
        total = sum(_rup(int(torch.tensor(shapes[n]).prod()), 64) for n in _NAMES) * nb
        f32 = dict(dtype=torch.float32, device=dev)
        self.master = torch.zeros(total, **f32)
        self.grad = torch.zeros(total, **f32)
        self.exp_avg = torch.zeros(total, **f32)
        self.exp_avg_sq = torch.zeros(total, **f32)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.P, self.Gd = [], []          # per block: name -> view
        off = 0
        # every block owns the same span of the flat buffers: [nb, ...] strided views address one tensor of all blocks
        numel = {n: int(torch.tensor(shapes[n]).prod()) for n in _NAMES}
        span = sum(_rup(numel[n], 64) for n in _NAMES)
        first, o = {}, 0
        for n in _NAMES:
            first[n] = o
            o += _rup(numel[n], 64)

        def all_blocks(flat, n):
            return torch.as_strided(flat, (nb,) + tuple(shapes[n]), (span,) + tuple(torch.empty(shapes[n]).stride()), first[n])
        self._w1_all, self._w2_all = all_blocks(self.master, "subnet.0.weight"), all_blocks(self.master, "subnet.2.weight")
        self._dgs_all, self._dgo_all = all_blocks(self.grad, "global_scale"), all_blocks(self.grad, "global_offset")
        for k in range(nb):
            pv, gv = {}, {}
            for n in _NAMES:
                numel = int(torch.tensor(shapes[n]).prod())
                pv[n] = self.master[off:off + numel].view(shapes[n])
                gv[n] = self.grad[off:off + numel].view(shapes[n])
                pv[n].copy_(params["module_list.%d.%s" % (k, n)].to(dev, torch.float32))
                off += _rup(numel, 64)
            self.P.append(pv)
            self.Gd.append(gv)
        self.fixed = {"module_list.%d.%s" % (k, n): params["module_list.%d.%s" % (k, n)].to(dev, torch.float32).contiguous()
                      for k in range(nb) for n in ("w_perm", "w_perm_inv")}
        self.flow = FlowPacked(C_dim, self.state_dict(), n_blocks=nb, device=dev)
        # ---- bf16 operands / workspaces of the parameter-gradient GEMMs
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.W1b = torch.zeros(nb, HIDDEN, 64, **bf)                    # K-major B of the hidden recompute
        self.W2b = torch.zeros(nb, _rup(2 * c2, 8), HIDDEN, **bf)       # MN-major B of the dgrad
        self.X1 = torch.zeros(nb, M, 64, **bf)                          # exports (padding stays zero)
        self.DS = torch.zeros(nb, M, 64, **bf)
        self.H = torch.empty(nb, M, HIDDEN, **bf)                       # relu(x1 W1^T + b1)
        self.DP = torch.empty(nb, M, HIDDEN, **bf)                      # (d_sub . W2) * relu'
        self.x = torch.zeros(batch, C_dim, **f32)
        self.noise = torch.zeros(batch, C_dim, **f32)
        self.u = torch.zeros(M, C_dim, **f32)
        self.loss = torch.zeros(1, **f32)
        # staging of the global-affine gradients (atomically accumulated by the kernel, [n_blocks, C] contiguous each) and
        # the NLL sum: ONE buffer, cleared by one launch per step
        self._acc = torch.zeros(2 * nb * C_dim + 1, **f32)
        self._dgs = self._acc[:nb * C_dim].view(nb, 1, C_dim)
        self._dgo = self._acc[nb * C_dim:2 * nb * C_dim].view(nb, 1, C_dim)
        self.nll_sum = self._acc[2 * nb * C_dim:]
        self._refresh_shadows()
        self._plans = None

    # ------------------------------------------------------------------------------------------
    def state_dict(self):
        sd = {}
        for k in range(self.nb):
            for n in _NAMES:
                sd["module_list.%d.%s" % (k, n)] = self.P[k][n]
        sd.update(self.fixed)
        return sd

    def _refresh_shadows(self):
        with torch.no_grad():                                   # all blocks at once: two cast launches
            self.W1b[:, :, :self.c1].copy_(self._w1_all)
            self.W2b[:, :2 * self.c2].copy_(self._w2_all)
        self.flow.repack(self.state_dict())

    def _prob(self, A, B, M, N, K, lda, ldb, **kw):
        P = GemmProblem()
        P.A, P.B, P.M, P.N, P.K, P.lda, P.ldb = A.data_ptr(), B.data_ptr(), M, N, K, lda, ldb
        P.flags = kw.pop("flags", 0)
        for name, t in kw.items():
            if name == "bias":
                P.bias = t.data_ptr()
            elif name == "out_f32":
                P.out_f32, P.ld_f32 = t.data_ptr(), t.stride(0)
            else:
                setattr(P, name, t.data_ptr())
                setattr(P, "ld_" + name, t.stride(0))
        return P

    def _build_plans(self):
        nb, M, c1, c2 = self.nb, self.M, self.c1, self.c2
        g1 = [self._prob(self.X1[k], self.W1b[k], M, HIDDEN, 64, 64, 64, flags=EPI_RELU_PRE, bias=self.P[k]["subnet.0.bias"],
                         out=self.H[k]) for k in range(nb)]
        g2 = [self._prob(self.DS[k], self.W2b[k], M, HIDDEN, 2 * c2, 64, HIDDEN, flags=GEMM_B_MN | EPI_YMASK_ZERO,
                         ymask=self.H[k], out=self.DP[k]) for k in range(nb)]
        g3 = [self._prob(self.DS[k], self.H[k], 2 * c2, HIDDEN, M, 64, HIDDEN, flags=GEMM_A_MN | GEMM_B_MN,
                         out_f32=self.Gd[k]["subnet.2.weight"]) for k in range(nb)]
        g4 = [self._prob(self.DP[k], self.X1[k], HIDDEN, c1, M, HIDDEN, 64, flags=GEMM_A_MN | GEMM_B_MN,
                         out_f32=self.Gd[k]["subnet.0.weight"]) for k in range(nb)]
        gemms = [((GemmProblem * nb)(*g), nb) for g in (g1, g2, g3, g4)]
        items = []
        for k in range(nb):
            items.append((self.DP[k].data_ptr(), HIDDEN, M, HIDDEN, self.Gd[k]["subnet.0.bias"].data_ptr()))
            items.append((self.DS[k].data_ptr(), 64, M, 2 * c2, self.Gd[k]["subnet.2.bias"].data_ptr()))
        arr = (_cabi.ColsumItem * len(items))()
        for j, (g, ldg, Mi, Ni, out) in enumerate(items):
            arr[j].G, arr[j].out, arr[j].ldg, arr[j].M, arr[j].N, arr[j].accumulate = g, out, ldg, Mi, Ni, 0
        self._plans = (gemms, (arr, len(items)))

    # ------------------------------------------------------------------------------------------
    def forward_backward(self):
        L = self.lib
        st = torch.cuda.current_stream().cuda_stream
        if self._plans is None:
            self._build_plans()
        if not self.external_rows:
            self.flow.sample(self.x, self.noise, self.u)           # [x ; s], s detached with the root joint zeroed
        # the NLL sum and the global-affine gradients are accumulated atomically into contiguous [n_blocks, C] staging
        # (their slots of the flat gradient buffer sit one block span apart): one clear before, two strided copies after
        self._acc.zero_()
        check(L.links_flow_nll_train(self.flow.packed.data_ptr(), self.C, self.nb, self.u.data_ptr(), self.M, 1.0 / self.B,
                                     self.nll_sum.data_ptr(), None, self.X1.data_ptr(), self.DS.data_ptr(),
                                     self._dgs.data_ptr(), self._dgo.data_ptr(), self.flow.stash_for(self.M).data_ptr(), st),
              "links_flow_nll_train")
        self._param_grads()
        torch.mul(self.nll_sum, 1.0 / self.B, out=self.loss)        # dist_2d + dist_2d_sample

    def _param_grads(self):
        """Exports of the fused kernel (X1, DS, global-affine sums) -> every parameter gradient in self.grad / self.Gd."""
        L = self.lib
        st = torch.cuda.current_stream().cuda_stream
        gemms, (carr, cn) = self._plans
        for arr, n in gemms:
            check(L.links_gemm_grouped(arr, n, st), "links_gemm_grouped")
        check(L.links_colsum_bf16_batched(carr, cn, st), "links_colsum_bf16_batched")
        self._dgs_all.copy_(self._dgs)
        self._dgo_all.copy_(self._dgo)

    def vjp(self, x, gz, gld=None, dx=None):
        """(z, log_jac_det) = inn(x) differentiated for arbitrary seeds: dx = J^T (gz, gld) (optional) and the gradient of
        every trainable parameter (-> self.Gd[k][name]), one fused kernel launch + the parameter-gradient GEMMs.
        x, gz: contiguous fp32 [rows, C]; gld: fp32 [rows] or None.  Backs autograd of the FrEIA shim when a flow is
        trained through `inn(x)` (train_full_pose_norm_flow.py:75-98)."""
        assert x.shape == (self.M, self.C) and gz.shape == x.shape and x.is_contiguous() and gz.is_contiguous()
        if self._plans is None:
            self._build_plans()
        self._acc.zero_()
        check(self.lib.links_flow_vjp_train(self.flow.packed.data_ptr(), self.C, self.nb, x.data_ptr(), self.M, gz.data_ptr(),
                                            gld.data_ptr() if gld is not None else None,
                                            dx.data_ptr() if dx is not None else None, self.X1.data_ptr(),
                                            self.DS.data_ptr(), self._dgs.data_ptr(), self._dgo.data_ptr(),
                                            self.flow.stash_for(self.M).data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "links_flow_vjp_train")
        self._param_grads()
        return dx

    def optimizer_step(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)
        st = torch.cuda.current_stream().cuda_stream
        check(self.lib.links_adam_step(self.master.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.master.numel(), self.lr, 0.9, 0.999, 1e-8, self.wd, 0,
                                       self.step_dev.data_ptr(), 1.0 / self.world, self.lr_dev.data_ptr(), st), "links_adam_step")
        self._refresh_shadows()

    def step(self):
        self.forward_backward()
        self.optimizer_step()

    def _graph_ok(self):
        return self.world == 1          # data-parallel: the gradient all-reduce stays an eager NCCL call

    def set_lr(self, lr):
        self.lr = lr
        self.lr_dev.fill_(float(lr))          # read on the device by the Adam kernel (graph replays follow the scheduler)

    def loss_dict(self):
        return {"loss": self.loss.item()}


class PartFlowTrainer(_GraphReplay):
    """Joint training step of the four part flows (reference train_leg_torso_left_right_norm_flow.py:100-198).

        parts of the data              -> NLL under the leg / torso / left / right flows           (:108-127)
        s = full_inn^-1(add_noise(full_inn(x))) (frozen full-pose flow, no grad), root zeroed       (:130-141)
        parts of s                     -> NLL under the same four flows                            (:144-161)
        loss = sum of the eight means; four independent Adam(wd 1e-5) steps                         (:163-174)

    The four flows share nothing but their input rows, so each runs its FlowTrainStep (one fused tensor-core launch over
    the stacked rows [data ; samples] + the parameter-gradient GEMMs + Adam) on its own stream."""

    NAMES = ("legs", "torso", "left", "right")

    def __init__(self, full_params, part_params, batch, lr=2e-4, weight_decay=1e-5, device="cuda", process_group=None):
        """full_params: frozen 34-d sampler; part_params: dict NAMES -> FrEIA-layout state dict."""
        from . import maps
        self.B = batch
        self.device = dev = torch.device(device)
        self.sampler = FlowPacked(34, full_params, device=dev)
        joints = {"legs": maps.LEG_JOINTS, "torso": maps.TORSO_JOINTS, "left": maps.LEFT_JOINTS, "right": maps.RIGHT_JOINTS}
        self.index = {n: torch.tensor(maps.part_index(joints[n]), dtype=torch.long, device=dev) for n in self.NAMES}
        self.steps = {n: FlowTrainStep(2 * len(joints[n]), part_params[n], batch, lr=lr, weight_decay=weight_decay, device=dev,
                                       process_group=process_group, external_rows=True) for n in self.NAMES}
        self.streams = {n: torch.cuda.Stream(device=dev) for n in self.NAMES}
        self.x = torch.zeros(batch, 34, dtype=torch.float32, device=dev)
        self.noise = torch.zeros(batch, 34, dtype=torch.float32, device=dev)
        self.u = torch.zeros(2 * batch, 34, dtype=torch.float32, device=dev)

    def step(self):
        main = torch.cuda.current_stream()
        self.sampler.sample(self.x, self.noise, self.u)            # [x ; s]
        for n in self.NAMES:
            st, side = self.steps[n], self.streams[n]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                torch.index_select(self.u, 1, self.index[n], out=st.u)     # integer gather (utils/helpers.py:55-65)
                st.step()
        for n in self.NAMES:
            main.wait_stream(self.streams[n])

    def _graph_ok(self):
        return all(st.world == 1 for st in self.steps.values())

    def set_lr(self, lr):
        for st in self.steps.values():
            st.set_lr(lr)

    def loss_dict(self):
        d = {"dist_2d_" + n: self.steps[n].loss.item() for n in self.NAMES}      # data + sample NLL of each flow
        d["loss"] = sum(d.values())
        return d

    def state_dict(self, name):
        return self.steps[name].state_dict()
