"""Occlusion-model training step (reference train_occlusion_models.py:144-314) and sharded evaluation
(eval_h36m.py:50-97) built from C-ABI launches only.

Occlusion step: frozen leg/torso lifters (pose branch only; the reference also evaluates the left/right lifters
:160-161 and never uses them) -> root-centred 3D pose (no depth clamp, :164-174) -> three rounds (identity, Ry, Ry.Ry)
of {integer gathers -> 8 predictors -> squared-error losses}; the three rounds are three passes of one 8-network
MlpSet, so every weight gradient is a single GEMM contracting over all rounds.
"""
import torch

from . import _cabi, maps
from ._cabi import check
from .mlp import MlpSet

OCC_NAMES = maps.OCC_NAMES
# predictor input / output widths (train_occlusion_models.py:90-97; utils/models_def.py:243-327)
OCC_IN = {n: (len(maps.occ_input_index(n)[0]) // maps.occ_input_index(n)[1]) for n in OCC_NAMES}
OCC_OUT = {n: len(maps.occ_target_index(n)) for n in OCC_NAMES}


class OcclusionStep:
    def __init__(self, batch, lifter_params, predictor_params, cfg=None, device="cuda", process_group=None,
                 comm_stream=None):
        """lifter_params: [leg, torso] state dicts; predictor_params: dict OCC_NAMES -> state dict."""
        if batch % 2:
            raise ValueError("split_data_left_right_3d (utils/helpers.py:81-91) needs an even batch")
        self.cfg = dict(depth=10.0, lr=2e-4, weight_decay=1e-5)
        self.cfg.update(cfg or {})
        self.B = batch
        self.device = torch.device(device)
        self.lib = _cabi.lib()
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        dev = self.device
        self.lifters = MlpSet("lifter", [14, 20], [{"downscale": 7, "angles": 1}, {"downscale": 10, "angles": 1}], batch,
                              n_passes=1, device=dev, train=False, pass_branches=[["pose"]])
        self.lifters.load_state_dicts(lifter_params)
        self.mlp = MlpSet("predictor", [OCC_IN[n] for n in OCC_NAMES], [{"downscale": OCC_OUT[n]} for n in OCC_NAMES],
                          batch, n_passes=3, device=dev, train=True)
        self.mlp.load_state_dicts([predictor_params[n] for n in OCC_NAMES])
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.x = torch.zeros(batch, 34, **f32)
        self.u_y = [torch.zeros(batch, **f32) for _ in range(2)]
        self.pose = [torch.zeros(batch, 51, **f32) for _ in range(3)]
        self.idx_lift = [torch.tensor(maps.part_index(j), **i32) for j in (maps.LEG_JOINTS, maps.TORSO_JOINTS)]
        self.idx_in, self.period = [], []
        for n in OCC_NAMES:
            idx, per = maps.occ_input_index(n)
            self.idx_in.append(torch.tensor(idx, **i32))
            self.period.append(per)
        self.idx_tgt = [torch.tensor(maps.occ_target_index(n), **i32) for n in OCC_NAMES]
        self.loss_sums = torch.zeros(8, **f32)
        self.losses = torch.zeros(9, **f32)
        self.comm = comm_stream if comm_stream is not None else torch.cuda.Stream(device=dev)

    def _on_bucket(self, b):
        main = torch.cuda.current_stream()
        self.comm.wait_stream(main)
        m = self.mlp
        with torch.cuda.stream(self.comm):
            if self.world > 1:
                if self.cfg.get("grad_comm", "bf16") == "bf16":
                    torch.distributed.all_reduce(m.compress_grads(b), group=self.pg)     # half the NVLink bytes
                else:
                    a, e = m.bucket_ranges[b]
                    torch.distributed.all_reduce(m.grad[a:e], group=self.pg)
            m.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world, bucket=b,
                        last=(b == len(m.buckets) - 1),
                        grads_bf16=(self.world > 1 and self.cfg.get("grad_comm", "bf16") == "bf16"),
                        rest_only=self._fuse_adam)

    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def forward_backward(self, fused_optimizer=False):
        L, B, m, lf = self.lib, self.B, self.mlp, self.lifters
        st = self._st()
        self._fuse_adam = bool(fused_optimizer and self.world == 1 and self.cfg.get("fuse_adam", True))
        if self._fuse_adam:
            m.adam_prepare(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0)
        for s in range(2):
            nidx = self.idx_lift[s].numel()
            check(L.links_pack_rows(self.x.data_ptr(), 34, B, self.idx_lift[s].data_ptr(), nidx, 1,
                                    lf.x0[0][s].data_ptr(), None, 0, 0, st), "links_pack_rows")
        lf.run(lf.forward_ops(0))
        check(L.links_occ_lift(self.x.data_ptr(), lf.head_out[0][0]["downscale"].data_ptr(),
                               lf.head_out[0][1]["downscale"].data_ptr(), B, self.cfg["depth"], self.pose[0].data_ptr(), st),
              "links_occ_lift")
        for r in (1, 2):
            check(L.links_occ_rotate_y(self.pose[r - 1].data_ptr(), self.u_y[r - 1].data_ptr(), B,
                                       self.pose[r].data_ptr(), st), "links_occ_rotate_y")
        self.loss_sums.zero_()
        for r in range(3):
            for s in range(8):
                n_idx = self.idx_in[s].numel() // self.period[s]
                check(L.links_pack_rows(self.pose[r].data_ptr(), 51, B, self.idx_in[s].data_ptr(), n_idx, self.period[s],
                                        m.x0[r][s].data_ptr(), None, 0, 0, st),
                      "links_pack_rows")
            m.run(m.forward_ops(r))
            for s in range(8):
                n_out = self.idx_tgt[s].numel()
                pred = m.head_out[r][s]["downscale"]
                check(L.links_occ_mse(pred.data_ptr(), pred.stride(0), self.pose[r].data_ptr(), self.idx_tgt[s].data_ptr(),
                                      n_out, B, 1.0 / B, self.loss_sums[s:s + 1].data_ptr(),
                                      m.G[r][s]["downscale"].data_ptr(), None, 0, 0, st), "links_occ_mse")
        for r in range(2):
            m.run(m.backward_ops(r, need_input_grad=False))
        m.run(m.backward_ops(2, need_input_grad=False, wgrad=True, split_at_buckets=self.world > 1, fuse_adam=self._fuse_adam),
              on_bucket=self._on_bucket if fused_optimizer else None)
        if fused_optimizer:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.losses[:8] = self.loss_sums / B
        self.losses[8] = self.losses[:8].sum()

    def optimizer_step(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.mlp.grad, group=self.pg)
        self.mlp.adam_step(lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"], grad_scale=1.0 / self.world)

    def step(self):
        self.forward_backward(fused_optimizer=True)

    def set_lr(self, lr):
        self.cfg["lr"] = lr
        self.mlp.set_lr(lr)           # device word read by the Adam kernel: captured graphs follow the schedule

    def capture(self, warmup=2):
        """Capture step() into one CUDA graph (replay with the returned graph's .replay())."""
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

    def loss_dict(self):
        v = self.losses.tolist()
        d = {"threed_loss_" + n: v[i] for i, n in enumerate(OCC_NAMES)}
        d["loss"] = v[8]
        return d


class EvalRunner:
    """Sharded evaluation: lift with the left/right (or leg/torso) lifters and score N-MPJPE, PA-MPJPE ('best', the
    number eval_h36m.py:83-99 prints) and the batched PA variant; per-rank double sums, one final reduction."""

    def __init__(self, kind, lifter_params, chunk=65536, depth=10.0, choice="right", device="cuda", process_group=None):
        self.kind, self.chunk, self.depth = kind, chunk, depth
        self.device = torch.device(device)
        self.lib = _cabi.lib()
        self.pg = process_group
        self.joints = maps.part_joint_lists(kind)
        nj = [len(j) for j in self.joints]
        self.nj = nj
        self.mlp = MlpSet("lifter", [2 * n for n in nj], [{"downscale": n, "angles": 1} for n in nj], chunk, n_passes=1,
                          device=device, train=False, pass_branches=[["pose"]])
        self.mlp.load_state_dicts(lifter_params)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.idx = [torch.tensor(maps.part_index(j), **i32) for j in self.joints]
        # combine map (utils/helpers.py:40-53): depth of full-pose joint j = head[src][col]; joint 0 forced to 0
        m = maps.geom_maps(kind)
        v = 0 if (kind == "lt" or choice == "left") else 1
        src = torch.tensor([m.src_net[v][j] for j in range(17)], **i32)
        col = torch.tensor([m.col[j] for j in range(17)], **i32)
        self.gather = (src.long() * 32 + col.long())
        self.depth_off = torch.zeros(chunk, 32, dtype=torch.float32, device=self.device)
        self.sums = torch.zeros(3, dtype=torch.float64, device=self.device)
        self.count = 0

    def reset(self):
        self.sums.zero_()
        self.count = 0

    def load_lifters(self, lifter_params):
        """Swap in new lifter weights (epoch-end validation of a model that is still training)."""
        self.mlp.load_state_dicts(lifter_params)

    def run_chunk(self, poses_2d, gt_3d, want_pred=False):
        """poses_2d [n,34], gt_3d [n,51] device fp32 tensors, n <= chunk.  want_pred: also return the lifted poses
        [n,51] (eval_h36m.py:60-66; validation PCK / AUC need them)."""
        n = poses_2d.shape[0]
        L, m = self.lib, self.mlp
        st = torch.cuda.current_stream().cuda_stream
        for s in range(2):
            check(L.links_pack_rows(poses_2d.data_ptr(), 34, n, self.idx[s].data_ptr(), 2 * self.nj[s], 1,
                                    m.x0[0][s].data_ptr(), None, 0, 0, st), "links_pack_rows")
        m.run(m.forward_ops(0, rows=n if n != m.M else None))
        # assemble the 17 depth offsets (integer gather of the two heads; root joint zeroed, eval_h36m.py:55-58)
        heads = torch.cat((m.head_out[0][0]["downscale"][:n], m.head_out[0][1]["downscale"][:n]), dim=1)
        d = heads[:, self.gather]
        d[:, 0] = 0.0
        self.depth_off[:n, :17] = d
        check(L.links_eval_lift_score(poses_2d.data_ptr(), self.depth_off.data_ptr(), 32, gt_3d.data_ptr(), n, self.depth,
                                      self.sums.data_ptr(), st), "links_eval_lift_score")
        self.count += n
        if want_pred:
            dd = (d + self.depth)
            return torch.cat((poses_2d[:, :17] * dd, poses_2d[:, 17:] * dd, dd), dim=1)
        return None

    def result(self):
        """Single final reduction over ranks (per-rank double sums + counts)."""
        from .shard import reduce_eval_sums
        s, total = reduce_eval_sums(self.sums.clone(), self.count, self.pg) if self.pg is not None else \
            ((self.sums / max(self.count, 1)).tolist(), self.count)
        return {"n_mpjpe": s[0], "pa_mpjpe": s[1], "pa_mpjpe_batch": s[2], "count": total}
