"""Step drivers: the self-supervised lifter training step (leg/torso and left/right), built only from
C-ABI kernel launches on preallocated buffers (CUDA-graph capturable; no host sync inside a step).

Reference step bodies: train_leg_torso_lifter.py:123-276, train_left_right_lifter.py:121-427.
Launch order of one step:
  sample (full-pose flow fwd -> noise -> rev)            flow_sample
  pack part inputs, lifter pass 1 (pose + angle)         pack_rows, grouped GEMMs
  elevation stats, lift/rotate/project                   elev_stats, geom_forward
  part-flow NLL forward + input gradient                 flow_nll_fwdbwd
  lifter pass 2 (pose branch only)                       pack_rows, grouped GEMMs
  consistency losses + d/d(pass-2 heads)                 geom_loss
  pass-2 dgrad chain -> d/d(projected 2D)                grouped GEMMs
  geometry backward (+ batch-statistic terms)            geom_backward, geom_backward_angles
  pass-1 dgrad chain, all wgrads, bias grads             grouped GEMMs, colsum
  [gradient all-reduce], Adam + bf16 shadow refresh      NCCL, adam_step, cast_weight
"""
import ctypes as C

import torch

from . import _cabi, maps
from ._cabi import check
from .flowpack import FlowPacked
from .mlp import MlpSet

DEFAULT_CFG = dict(depth=10.0, weight_bl=50.0, weight_2d=1.0, weight_3d=1.0, weight_likeli=1.0, weight_velocity=1.0,
                   lr=2e-4, weight_decay=1e-5)


class LifterStep:
    """kind = 'lt' (Leg_Lifter + Torso_Lifter) or 'lr' (left + right Left_Right_Lifter)."""

    def __init__(self, kind, batch, lifter_params, part_flow_params, full_flow_params, cfg=None, device="cuda",
                 process_group=None, comm_stream=None):
        self.kind = kind
        self.cfg = dict(DEFAULT_CFG)
        self.cfg.update(cfg or {})
        self.B = batch
        self.N = 2 * batch
        self.device = torch.device(device)
        self.lib = _cabi.lib()
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        dev = self.device
        self.maps = maps.geom_maps(kind, self.cfg)
        self.joints = maps.part_joint_lists(kind)
        nj = [len(j) for j in self.joints]
        self.nj = nj
        self.mlp = MlpSet("lifter", [2 * n for n in nj], [{"downscale": n, "angles": 1} for n in nj], self.N,
                          n_passes=2, device=dev, train=True, pass_branches=[["pose", "angle"], ["pose"]],
                          max_buckets=self.cfg.get("dp_buckets") if (self.world > 1 or self.cfg.get("dp_layout")) else None)
        self.mlp.load_state_dicts(lifter_params)
        self.part_flows = [FlowPacked(2 * nj[s], part_flow_params[s], device=dev) for s in range(2)]
        self.full_flow = FlowPacked(34, full_flow_params, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        N, B = self.N, self.B
        # inputs (filled by the caller before step())
        self.x = torch.zeros(B, 34, **f32)
        self.noise = torch.zeros(B, 34, **f32)
        self.eps_x = torch.zeros(N, **f32)
        self.u_y = torch.zeros(N, **f32)
        # intermediates
        self.u = torch.zeros(N, 34, **f32)
        self.stats = torch.zeros(2, **f32)
        self.qpart = [torch.zeros(N, 2 * nj[s], **f32) for s in range(2)]
        self.qfull = [torch.zeros(N, 34, **f32) for _ in range(2)]
        self.dflow = [torch.zeros(N, 2 * nj[s], **f32) for s in range(2)]
        self.scal = torch.zeros(8, **f32)     # [0:4] loss sums (L3d, rep, pair, bl), [4:6] nll sums, [6:8] red
        self.dgamma = torch.zeros(N, **f32)
        self.da = torch.zeros(N, **f32)
        self.losses = torch.zeros(8, **f32)   # L3d, rep_rot, re_rot_3d, bl_prior, likeli_0, likeli_1, likeli, loss
        i32 = dict(dtype=torch.int32, device=dev)
        self.idx_u = [torch.tensor(maps.part_index(self.joints[s]), **i32) for s in range(2)]
        self.idx_q = [torch.arange(2 * nj[s], **i32) for s in range(2)]
        self._norm = torch.tensor([1.0 / N, 1.0 / N, 1.0 / max(N // 2, 1), 1.0 / N, 1.0 / N, 1.0 / N], **f32)
        c = self.cfg
        self._w = torch.tensor([c["weight_3d"], c["weight_2d"], c["weight_velocity"], c["weight_bl"],
                                c["weight_likeli"], c["weight_likeli"]], **f32)
        self.graph = None
        # the two part-flow NLL kernels (few CTAs each, ~0.3 ms) run on forked streams next to the pass-2 GEMMs;
        # they are joined right before the geometry backward that consumes their input gradients
        self._flow_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        # gradient buckets are all-reduced (NCCL), Adam-updated and re-cast on this stream as soon as the backward
        # pass has produced them, overlapping the rest of backward.  Steps that run concurrently (StepGroup) must
        # share ONE comm stream so that every rank issues its collectives in the same order.
        self.comm = comm_stream if comm_stream is not None else torch.cuda.Stream(device=dev)
        # Adam + shadow refresh of a reduced bucket run on their own stream so the collectives stay back to back
        self.opt_stream = torch.cuda.Stream(device=dev)

    # ------------------------------------------------------------------------------------------
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def _pack(self, src, idx, n_idx, p, s):
        m = self.mlp
        check(self.lib.links_pack_rows(src.data_ptr(), src.stride(0), self.N, idx.data_ptr(), n_idx, 1,
                                       m.x0[p][s].data_ptr(), None, 0, 0, self._st()),
              "links_pack_rows")

    def _on_bucket(self, b):
        """Bucket b of the flat gradient buffer is final: all-reduce + Adam + shadow refresh on the comm stream."""
        main = torch.cuda.current_stream()
        m = self.mlp
        src = main
        if self.world > 1:
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                if self.cfg.get("grad_comm", "bf16") == "bf16":
                    torch.distributed.all_reduce(m.compress_grads(b), group=self.pg)     # half the NVLink bytes
                else:
                    a, e = m.bucket_ranges[b]
                    torch.distributed.all_reduce(m.grad[a:e], group=self.pg)
            src = self.comm
        self.opt_stream.wait_stream(src)
        with torch.cuda.stream(self.opt_stream):
            m.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world, bucket=b,
                        last=(b == len(m.buckets) - 1),
                        grads_bf16=(self.world > 1 and self.cfg.get("grad_comm", "bf16") == "bf16"))

    def forward_backward(self, fused_optimizer=False):
        """Everything of one step up to (and including) the gradients.  fused_optimizer=True also runs the
        per-bucket all-reduce / Adam / shadow refresh overlapped with the backward pass (what step() does)."""
        L, m, N = self.lib, self.mlp, self.N
        mp = C.byref(self.maps)
        self.full_flow.sample(self.x, self.noise, self.u)
        for s in range(2):
            self._pack(self.u, self.idx_u[s], 2 * self.nj[s], 0, s)
        m.run(m.forward_plan(0))
        h1 = [m.head_out[0][s]["downscale"] for s in range(2)]
        a1 = [m.head_out[0][s]["angles"] for s in range(2)]
        h2 = [m.head_out[1][s]["downscale"] for s in range(2)]
        check(L.links_elev_stats(a1[0].data_ptr(), a1[1].data_ptr(), N, self.stats.data_ptr(), self._st()),
              "links_elev_stats")
        common = [self.u.data_ptr(), h1[0].data_ptr(), h1[1].data_ptr(), a1[0].data_ptr(), a1[1].data_ptr(),
                  self.eps_x.data_ptr(), self.u_y.data_ptr(), self.stats.data_ptr()]
        check(L.links_geom_forward(mp, *common, N, self.qpart[0].data_ptr(), self.qpart[1].data_ptr(),
                                   self.qfull[0].data_ptr(), self.qfull[1].data_ptr(), self._st()), "links_geom_forward")
        self.scal.zero_()
        main = torch.cuda.current_stream()
        for s in range(2):
            fs = self._flow_streams[s]
            fs.wait_stream(main)
            with torch.cuda.stream(fs):
                self.part_flows[s].nll_fwdbwd(self.qpart[s], self.cfg["weight_likeli"] / N, self.scal[4 + s:5 + s],
                                              self.dflow[s])
        for s in range(2):
            self._pack(self.qpart[s], self.idx_q[s], 2 * self.nj[s], 1, s)
        m.run(m.forward_plan(1))
        g2 = [m.G[1][s]["downscale"] for s in range(2)]
        check(L.links_geom_loss(mp, *common, h2[0].data_ptr(), h2[1].data_ptr(), N, self.scal.data_ptr(),
                                g2[0].data_ptr(), g2[1].data_ptr(), None, None, 0, 0, self._st()), "links_geom_loss")
        m.run(m.backward_plan(1, need_input_grad=True))
        g1 = [m.G[0][s]["downscale"] for s in range(2)]
        ga = [m.G[0][s]["angles"] for s in range(2)]
        for fs in self._flow_streams:
            main.wait_stream(fs)
        check(L.links_geom_backward(mp, *common, h2[0].data_ptr(), h2[1].data_ptr(), self.dflow[0].data_ptr(),
                                    self.dflow[1].data_ptr(), m.din[1][0].data_ptr(), m.din[1][1].data_ptr(), N,
                                    g1[0].data_ptr(), g1[1].data_ptr(), None, None, 0, 0,
                                    self.dgamma.data_ptr(), self.da.data_ptr(), self.scal[6:8].data_ptr(), self._st()),
              "links_geom_backward")
        check(L.links_geom_backward_angles(a1[0].data_ptr(), a1[1].data_ptr(), self.eps_x.data_ptr(),
                                           self.stats.data_ptr(), self.dgamma.data_ptr(), self.scal[6:8].data_ptr(), N,
                                           ga[0].data_ptr(), ga[1].data_ptr(), None, None, 0, 0, self._st()),
              "links_geom_backward_angles")
        m.run(m.backward_plan(0, need_input_grad=False, wgrad=True), on_bucket=self._on_bucket if fused_optimizer else None)
        if fused_optimizer:
            main.wait_stream(self.opt_stream)
        # loss scalars (device side, no sync): L3d, rep_rot, re_rot_3d, bl_prior, likeli_0, likeli_1, likeli, loss
        t = self.scal[:6] * self._norm
        self.losses[:6] = t
        self.losses[6] = t[4] + t[5]
        self.losses[7] = (t * self._w).sum()

    def optimizer_step(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.mlp.grad, group=self.pg)
        self.mlp.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world)

    def step(self):
        self.forward_backward(fused_optimizer=True)

    def loss_dict(self):
        v = self.losses.tolist()
        names = ("L3d", "rep_rot", "re_rot_3d", "bl_prior")
        d = dict(zip(names, v[:4]))
        if self.kind == "lt":
            d["leg_likeli"], d["torso_likeli"] = v[4], v[5]
        else:
            d["likeli_right"], d["likeli_left"] = v[4], v[5]   # the reference swaps the names (:334-342)
        d["likeli"], d["loss"] = v[6], v[7]
        return d

    def set_lr(self, lr):
        self.cfg["lr"] = lr


class StepGroup:
    """Independent step objects (e.g. the leg/torso and the left/right lifter steps of one batch) run as parallel
    branches: step i is issued on its own stream forked from / joined to the caller's stream, so inside one captured
    CUDA graph the branches overlap -- one branch's few-CTA kernels (flows, geometry, heads) and launch gaps are
    filled by the other branch's GEMM tiles."""

    def __init__(self, steps):
        self.steps = list(steps)
        self.streams = [torch.cuda.Stream() for _ in self.steps[1:]]
        self.graph = None
        # Branches that share a communicator must share ONE comm stream: collectives are then issued (and captured) in
        # one deterministic order on every rank -- two NCCL kernels of one communicator racing on different streams
        # would deadlock.  Branches with their own process group (own communicator) keep their own comm stream and
        # their all-reduces run concurrently.
        by_pg = {}
        for st in self.steps:
            key = id(st.pg)
            if key in by_pg:
                st.comm = by_pg[key].comm
            else:
                by_pg[key] = st

    def step(self):
        main = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(main)
        for step, st in zip(self.steps[1:], self.streams):
            with torch.cuda.stream(st):
                step.step()
        self.steps[0].step()
        for st in self.streams:
            main.wait_stream(st)
        for step in self.steps:
            main.wait_stream(step.opt_stream)

    def capture(self, warmup=2):
        """Capture step() into one CUDA graph (replay with .replay()).  Eager warm-up runs first (lazy inits)."""
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

    def replay(self):
        self.graph.replay()
