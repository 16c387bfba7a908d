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


class _Kind:
    """Per-flavour state of a LifterStep: 'lt' (Leg_Lifter + Torso_Lifter) or 'lr' (left + right Left_Right_Lifter)."""
    pass


class LifterStep:
    """kind = 'lt' (Leg_Lifter + Torso_Lifter), 'lr' (left + right Left_Right_Lifter) or 'both'.

    'both' runs the leg/torso step and the left/right step of the SAME batch as one step: the four lifters are one
    4-network MlpSet, so every layer of both flavours sits in the same GEMM launches (twice the tiles per chain launch,
    one gradient bucket stream), the sampling flow is evaluated once, and only the small geometry / flow kernels are
    issued per flavour.  lifter_params / part_flow_params are then [leg, torso, left, right]."""

    def __init__(self, kind, batch, lifter_params, part_flow_params, full_flow_params, cfg=None, device="cuda",
                 process_group=None, comm_stream=None):
        self.kind = kind
        self.kinds = ["lt", "lr"] if kind == "both" else [kind]
        self.cfg = dict(DEFAULT_CFG)
        self.cfg.update(cfg or {})
        self.B = batch
        self.N = 2 * batch
        self.device = torch.device(device)
        self.lib = _cabi.lib()
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        N, B = self.N, self.B
        c = self.cfg
        self.K = []
        nj_all = []
        for ki, kd in enumerate(self.kinds):
            k = _Kind()
            k.kind, k.s0 = kd, 2 * ki
            k.maps = maps.geom_maps(kd, self.cfg)
            k.joints = maps.part_joint_lists(kd)
            k.nj = [len(j) for j in k.joints]
            nj_all += k.nj
            k.part_flows = [FlowPacked(2 * k.nj[s], part_flow_params[k.s0 + s], device=dev) for s in range(2)]
            k.stats = torch.zeros(2, **f32)
            k.qpart = [torch.zeros(N, 2 * k.nj[s], **f32) for s in range(2)]
            k.qfull = [torch.zeros(N, 34, **f32) for _ in range(2)]
            k.dflow = [torch.zeros(N, 2 * k.nj[s], **f32) for s in range(2)]
            k.scal = torch.zeros(8, **f32)     # [0:4] loss sums (L3d, rep, pair, bl), [4:6] nll sums, [6:8] red
            k.dgamma = torch.zeros(N, **f32)
            k.da = torch.zeros(N, **f32)
            k.losses = torch.zeros(8, **f32)   # L3d, rep_rot, re_rot_3d, bl_prior, likeli_0, likeli_1, likeli, loss
            k.idx_u = [torch.tensor(maps.part_index(k.joints[s]), **i32) for s in range(2)]
            k.idx_q = [torch.arange(2 * k.nj[s], **i32) for s in range(2)]
            # the two part-flow NLL kernels (few CTAs each) run on forked streams next to the pass-2 GEMMs; they are
            # joined right before the geometry backward that consumes their input gradients
            k.flow_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
            k.stream = torch.cuda.Stream(device=dev) if ki > 0 else None      # small per-flavour kernels of flavour > 0
            self.K.append(k)
        # data-parallel gradient exchange: "push" = reduce-scatter by peer stores fused into the wgrad epilogues + sharded
        # Adam + shadow all-gather by peer stores (mlp.MlpSet zero_group); "bf16" / "fp32" = bucketed NCCL all-reduce
        self.push = self.world > 1 and self.cfg.get("grad_comm", "bf16") == "push"
        if self.push:
            self.cfg["dp_buckets"] = 1            # no bucketed all-reduce: one contiguous "rest" range for the small layers
        self.mlp = MlpSet("lifter", [2 * n for n in nj_all], [{"downscale": n, "angles": 1} for n in nj_all], self.N,
                          n_passes=2, device=dev, train=True, pass_branches=[["pose", "angle"], ["pose"]],
                          zero_group=process_group if self.push else None,
                          head_groups=[[(k.s0, "downscale"), (k.s0 + 1, "downscale"), (k.s0, "angles"), (k.s0 + 1, "angles")]
                                       for k in self.K],
                          max_buckets=self.cfg.get("dp_buckets") if (self.world > 1 or self.cfg.get("dp_layout")) else None)
        self.mlp.load_state_dicts(lifter_params)
        self.full_flow = FlowPacked(34, full_flow_params, device=dev)
        # cfg global_elevation_stats: props.mean() / props.std() (train_leg_torso_lifter.py:168) over the GLOBAL batch -- two
        # tiny all-reduces per step (sums of gamma / gamma^2 forward, the two statistic cotangents backward) make the
        # data-parallel step equal to the reference's step on the concatenated batch; default: the rank's shard (DDP)
        self.gstats = bool(self.cfg.get("global_elevation_stats")) and self.world > 1
        self._esums = torch.zeros(2 * len(self.K), dtype=torch.float64, device=dev)
        self._red_all = torch.zeros(2 * len(self.K), **f32)
        for i, k in enumerate(self.K):
            k.esum = self._esums[2 * i:2 * i + 2]
            k.red = self._red_all[2 * i:2 * i + 2] if self.gstats else k.scal[6:8]
        # inputs (filled by the caller before step())
        self.x = torch.zeros(B, 34, **f32)
        self.noise = torch.zeros(B, 34, **f32)
        self.eps_x = torch.zeros(N, **f32)
        self.u_y = torch.zeros(N, **f32)
        self.u = torch.zeros(N, 34, **f32)
        # loss summary as one [8 x 8] mat-vec on the device-side sums scal[0:8] = (L3d, rep, pair, bl, nll_0, nll_1, -, -):
        # rows 0-5 the means, row 6 likeli = nll_0 + nll_1, row 7 the weighted total (train_leg_torso_lifter.py:266-272)
        norm = [1.0 / N, 1.0 / N, 1.0 / max(N // 2, 1), 1.0 / N, 1.0 / N, 1.0 / N]
        w = [c["weight_3d"], c["weight_2d"], c["weight_velocity"], c["weight_bl"], c["weight_likeli"], c["weight_likeli"]]
        lm = torch.zeros(8, 8, dtype=torch.float32)
        for i in range(6):
            lm[i, i] = norm[i]
            lm[7, i] = w[i] * norm[i]
        lm[6, 4], lm[6, 5] = norm[4], norm[5]
        self._loss_mat = lm.to(dev)
        self.graph = None
        # Sampling prefetch (cfg prefetch_sample): the sampling block (train_leg_torso_lifter.py:133-142) depends on no
        # trained weight, so the poses of step k+1 are drawn WHILE step k runs: step() first moves the prefetched
        # poses u_next -> u, then issues the next sampling pass on a side stream.  Protocol: before step k the caller
        # loads x / noise of step k+1 (and eps_x / u_y of step k); prime() draws the very first batch.
        self.prefetch = bool(self.cfg.get("prefetch_sample", False))
        # SM partition.  A chain launch is a persistent grid; kernels that should run NEXT to it need SMs of their own:
        #   window chains (forward pass 2, its dgrad chain) run beside the part-flow NLL kernels (ceil(N/128) CTAs each),
        #   the tail chain (pass-1 dgrad + weight gradients) runs beside the prefetched sampling pass (ceil(B/128) CTAs)
        #   and, data-parallel, beside the NCCL all-reduce CTAs.
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        tiles = (N + 127) // 128
        r_win = self.cfg.get("reserve_window")
        if r_win is None:
            # one SM per part-flow CTA while that is at most one wave on under half of the machine (B <= 1024: the flows finish
            # in one wave beside the window chains); for larger batches the flow CTAs come in many waves anyway and a small
            # reservation measured best (profiles/r02_sweep_reserve_*.txt: B = 2048 / 4096 / 8192 steps 8 % faster with 16)
            want = 2 * len(self.kinds) * tiles
            r_win = want if want <= n_sms // 2 - 10 else 16
        r_tail = self.cfg.get("reserve_tail")
        if r_tail is None:
            r_tail = (min((B + 127) // 128, 16) if self.prefetch else 0) + (int(self.cfg.get("nccl_ctas") or 0) if self.world > 1 else 0)
        self._ctas_window = (n_sms - r_win) // 2 * 2 if r_win > 0 else None
        self._ctas_tail = (n_sms - r_tail) // 2 * 2 if r_tail > 0 else None
        self.u_next = torch.zeros(N, 34, **f32) if self.prefetch else None
        self._sample_stream = torch.cuda.Stream(device=dev) if self.prefetch else None
        # gradient buckets are all-reduced (NCCL), Adam-updated and re-cast on this stream as soon as the backward
        # pass has produced them, overlapping the rest of backward.  Steps that run concurrently (StepGroup) must
        # share ONE comm stream so that every rank issues its collectives in the same order.
        self.comm = comm_stream if comm_stream is not None else torch.cuda.Stream(device=dev)
        # Adam + shadow refresh of a reduced bucket run on their own stream so the collectives stay back to back
        self.opt_stream = torch.cuda.Stream(device=dev)
        if len(self.K) == 1:       # single-flavour attribute names (tests / scripts)
            k = self.K[0]
            self.maps, self.joints, self.nj, self.part_flows = k.maps, k.joints, k.nj, k.part_flows
            self.stats, self.qpart, self.qfull, self.dflow, self.scal = k.stats, k.qpart, k.qfull, k.dflow, k.scal
            self.dgamma, self.da, self.losses = k.dgamma, k.da, k.losses

    # ------------------------------------------------------------------------------------------
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def _pack(self, src, idx, n_idx, p, s):
        m = self.mlp
        check(self.lib.links_pack_rows(src.data_ptr(), src.stride(0), self.N, idx.data_ptr(), n_idx, 1,
                                       m.x0[p][s].data_ptr(), None, 0, 0, self._st()),
              "links_pack_rows")

    def _on_bucket_push(self, b):
        """Push mode: the chain has stored every big weight gradient into its owner's staging buffer.  All-reduce the
        small rest (biases, upscale, heads: ~100 K values) over NCCL, then barrier -> sharded Adam (+ shadow stores into
        every rank) -> Adam of the rest -> barrier."""
        main = torch.cuda.current_stream()
        m = self.mlp
        a, e = m.bucket_mid[b], m.bucket_ranges[b][1]
        self.comm.wait_stream(main)
        with torch.cuda.stream(self.comm):
            torch.distributed.all_reduce(m.grad[a:e], group=self.pg)
        m.zero_barrier()                      # every rank's pushes have landed
        m.zero_adam()
        main.wait_stream(self.comm)
        m.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world, bucket=b,
                    last=(b == len(m.buckets) - 1), rest_only=True)
        m.zero_barrier()                      # every rank's new shadows have landed (and the staging slots are free again)

    def _on_bucket(self, b):
        """Bucket b of the flat gradient buffer is final: all-reduce + Adam + shadow refresh on the comm stream."""
        if self.push:
            return self._on_bucket_push(b)
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
                        grads_bf16=(self.world > 1 and self.cfg.get("grad_comm", "bf16") == "bf16"),
                        rest_only=self._fuse_adam)     # fused: the big layers were updated in the wgrad epilogues

    class _Fork:
        """`with step._fork(k):` -- the small kernels of flavour k > 0 go to that flavour's side stream (forked from the
        current stream at entry); _join() makes the main stream wait for them."""

        def __init__(self, k):
            self.k = k
            self.ctx = None

        def __enter__(self):
            if self.k.stream is not None:
                self.k.stream.wait_stream(torch.cuda.current_stream())
                self.ctx = torch.cuda.stream(self.k.stream)
                self.ctx.__enter__()

        def __exit__(self, *a):
            if self.ctx is not None:
                self.ctx.__exit__(*a)

    def _fork(self, k):
        return LifterStep._Fork(k)

    def _join(self):
        main = torch.cuda.current_stream()
        for k in self.K:
            if k.stream is not None:
                main.wait_stream(k.stream)

    def prime(self):
        """Draw the first batch of sampled poses from the current x / noise (sampling prefetch only)."""
        self.full_flow.sample(self.x, self.noise, self.u_next)

    def forward_backward(self, fused_optimizer=False):
        """Everything of one step up to (and including) the gradients.  fused_optimizer=True also runs the
        per-bucket all-reduce / Adam / shadow refresh overlapped with the backward pass (what step() does)."""
        L, m, N = self.lib, self.mlp, self.N
        main = torch.cuda.current_stream()
        # single GPU: no gradient leaves the device, so the optimiser step of the big layers runs inside the weight-gradient
        # epilogues (no stored gradients, no separate Adam / shadow-cast pass over 59 M parameters)
        self._fuse_adam = bool(fused_optimizer and self.world == 1 and self.cfg.get("fuse_adam", True))
        if self._fuse_adam:
            m.adam_prepare(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0)
        push = self.push and fused_optimizer
        if push:
            m.adam_prepare(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world)
        if self.prefetch:
            self.u.copy_(self.u_next)
        else:
            self.full_flow.sample(self.x, self.noise, self.u)
        for k in self.K:
            for s in range(2):
                self._pack(self.u, k.idx_u[s], 2 * k.nj[s], 0, k.s0 + s)
        m.run(m.forward_ops(0))
        if self.gstats:
            for k in self.K:
                a1 = [m.head_out[0][k.s0 + s]["angles"] for s in range(2)]
                check(L.links_elev_sums(a1[0].data_ptr(), a1[1].data_ptr(), N, k.esum.data_ptr(), self._st()), "links_elev_sums")
            torch.distributed.all_reduce(self._esums, group=self.pg)
            self._red_all.zero_()
        for k in self.K:
            with self._fork(k):
                mp = C.byref(k.maps)
                a1 = [m.head_out[0][k.s0 + s]["angles"] for s in range(2)]
                if self.gstats:
                    check(L.links_elev_finalize(k.esum.data_ptr(), N * self.world, k.stats.data_ptr(), self._st()),
                          "links_elev_finalize")
                else:
                    check(L.links_elev_stats(a1[0].data_ptr(), a1[1].data_ptr(), N, k.stats.data_ptr(), self._st()),
                          "links_elev_stats")
                k.common = [self.u.data_ptr()] + [m.head_out[0][k.s0 + s]["downscale"].data_ptr() for s in range(2)] + \
                           [a1[0].data_ptr(), a1[1].data_ptr(), self.eps_x.data_ptr(), self.u_y.data_ptr(), k.stats.data_ptr()]
                qf = [q.data_ptr() for q in k.qfull] if self.cfg.get("store_rot_2d", True) else [None, None]
                check(L.links_geom_forward(mp, *k.common, N, k.qpart[0].data_ptr(), k.qpart[1].data_ptr(), qf[0], qf[1],
                                           self._st()), "links_geom_forward")
                k.scal.zero_()
                here = torch.cuda.current_stream()
                for s in range(2):
                    fs = k.flow_streams[s]
                    fs.wait_stream(here)
                    with torch.cuda.stream(fs):
                        k.part_flows[s].nll_fwdbwd(k.qpart[s], self.cfg["weight_likeli"] / N, k.scal[4 + s:5 + s], k.dflow[s])
                for s in range(2):
                    self._pack(k.qpart[s], k.idx_q[s], 2 * k.nj[s], 1, k.s0 + s)
        self._join()
        m.run(m.forward_ops(1, max_ctas=self._ctas_window))
        for k in self.K:
            with self._fork(k):
                h2 = [m.head_out[1][k.s0 + s]["downscale"] for s in range(2)]
                g2 = [m.G[1][k.s0 + s]["downscale"] for s in range(2)]
                check(L.links_geom_loss(C.byref(k.maps), *k.common, h2[0].data_ptr(), h2[1].data_ptr(), N, k.scal.data_ptr(),
                                        g2[0].data_ptr(), g2[1].data_ptr(), None, None, 0, 0, self._st()), "links_geom_loss")
        self._join()
        m.run(m.backward_ops(1, need_input_grad=True, max_ctas=self._ctas_window))
        for k in self.K:
            with self._fork(k):
                here = torch.cuda.current_stream()
                for fs in k.flow_streams:
                    here.wait_stream(fs)
                h2 = [m.head_out[1][k.s0 + s]["downscale"] for s in range(2)]
                g1 = [m.G[0][k.s0 + s]["downscale"] for s in range(2)]
                ga = [m.G[0][k.s0 + s]["angles"] for s in range(2)]
                a1 = [m.head_out[0][k.s0 + s]["angles"] for s in range(2)]
                check(L.links_geom_backward(C.byref(k.maps), *k.common, h2[0].data_ptr(), h2[1].data_ptr(),
                                            k.dflow[0].data_ptr(), k.dflow[1].data_ptr(), m.din[1][k.s0].data_ptr(),
                                            m.din[1][k.s0 + 1].data_ptr(), N, g1[0].data_ptr(), g1[1].data_ptr(), None, None,
                                            0, 0, k.dgamma.data_ptr(), k.da.data_ptr(), k.red.data_ptr(), self._st()),
                      "links_geom_backward")
                if not self.gstats:
                    check(L.links_geom_backward_angles(a1[0].data_ptr(), a1[1].data_ptr(), self.eps_x.data_ptr(),
                                                       k.stats.data_ptr(), k.dgamma.data_ptr(), k.red.data_ptr(), N,
                                                       ga[0].data_ptr(), ga[1].data_ptr(), None, None, 0, 0, 0, self._st()),
                          "links_geom_backward_angles")
        self._join()
        if self.gstats:
            # the statistic's cotangents (sum da, sum eps * da) over ALL ranks' rows, then the angle-head gradients
            torch.distributed.all_reduce(self._red_all, group=self.pg)
            for k in self.K:
                a1 = [m.head_out[0][k.s0 + s]["angles"] for s in range(2)]
                ga = [m.G[0][k.s0 + s]["angles"] for s in range(2)]
                check(L.links_geom_backward_angles(a1[0].data_ptr(), a1[1].data_ptr(), self.eps_x.data_ptr(),
                                                   k.stats.data_ptr(), k.dgamma.data_ptr(), k.red.data_ptr(), N,
                                                   ga[0].data_ptr(), ga[1].data_ptr(), None, None, 0, 0, N * self.world, self._st()),
                      "links_geom_backward_angles")
        if self.prefetch:
            # poses of the NEXT step, drawn beside the longest GEMM chain of this one (which leaves it the SMs it needs)
            self._sample_stream.wait_stream(main)
            with torch.cuda.stream(self._sample_stream):
                self.full_flow.sample(self.x, self.noise, self.u_next)
        m.run(m.backward_ops(0, need_input_grad=False, wgrad=True, split_at_buckets=self.world > 1 and not push,
                             max_ctas=self._ctas_tail, fuse_adam=self._fuse_adam, push=push),
              on_bucket=self._on_bucket if fused_optimizer else None)
        if fused_optimizer and not push:
            main.wait_stream(self.opt_stream)       # (push mode runs its optimiser on the main stream)
        if self.prefetch:
            main.wait_stream(self._sample_stream)
        # loss scalars (device side, no sync): L3d, rep_rot, re_rot_3d, bl_prior, likeli_0, likeli_1, likeli, loss
        for k in self.K:
            check(L.links_small_matvec(self._loss_mat.data_ptr(), k.scal.data_ptr(), 8, 8, k.losses.data_ptr(), self._st()),
                  "links_small_matvec")

    def optimizer_step(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.mlp.grad, group=self.pg)
        self.mlp.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world)

    def step(self):
        self.forward_backward(fused_optimizer=True)

    @staticmethod
    def _loss_dict(k):
        v = k.losses.tolist()
        names = ("L3d", "rep_rot", "re_rot_3d", "bl_prior")
        d = dict(zip(names, v[:4]))
        if k.kind == "lt":
            d["leg_likeli"], d["torso_likeli"] = v[4], v[5]
        else:
            d["likeli_right"], d["likeli_left"] = v[4], v[5]   # the reference swaps the names (:334-342)
        d["likeli"], d["loss"] = v[6], v[7]
        return d

    def loss_dict(self):
        """Loss names of the reference's step; for kind='both' a dict {'lt': {...}, 'lr': {...}}."""
        if len(self.K) == 1:
            return self._loss_dict(self.K[0])
        return {k.kind: self._loss_dict(k) for k in self.K}

    def set_lr(self, lr):
        self.cfg["lr"] = lr
        self.mlp.set_lr(lr)           # device word read by the Adam kernel: captured graphs follow the schedule

    def capture(self, warmup=2):
        """Capture step() into one CUDA graph (replay with .replay()); eager warm-up runs first (lazy plan builds)."""
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
        from . import mlp as _mlp
        if _mlp.USE_CHAIN:
            # Chain launches are persistent kernels that spin on each other's tiles: two of them must never share the
            # machine (each could hold SMs the other's unscheduled clusters need).  Run the branches back to back.
            for step in self.steps:
                step.step()
            return
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
