"""Residual-MLP engine: S networks of identical topology run as grouped tcgen05 GEMM launches.

Replaces the forward and autograd-backward of reference utils/models_def.py (res_block :10-39, lifters
:111-239, occlusion predictors :243-327).  All device buffers are allocated once through PyTorch; every
operation is a C-ABI launch on the current stream, so a whole step is CUDA-graph capturable.

Data layout (per network s, per pass p):
  * fp32 master parameters / gradients / Adam moments live in flat buffers (one Adam launch, one all-reduce).
  * bf16 shadow weights per layer: W [N, Kp] -- the K-major B operand of forward and, read as an MN-major operand,
    the B operand of dgrad (no transposed shadow).
  * activations / gradients are bf16 row-major [n_passes * M, 1024]: pass p owns rows [p*M, (p+1)*M), so one wgrad
    GEMM per layer contracts over both passes, reading G and X as MN-major operands (no transposed copies).
  * per block a [M, 32]-word sign mask of the l2 pre-activation (needed exactly for leaky' in backward).
"""

import torch

from . import _cabi
import ctypes as C
import os

from ._cabi import EPI_LEAKY_POST, EPI_LEAKY_PRE, GEMM_A_MN, GEMM_B_MN, ChainPlan, ChainProblem, GemmProblem, HEAD_LD, check

# Chain launches (one persistent kernel per forward pass / dgrad chain, csrc/gemm.cu) are the default; LINKS_GEMM_CHAIN=0
# keeps the layer-by-layer grouped launches (A/B measurements).
USE_CHAIN = os.environ.get("LINKS_GEMM_CHAIN", "1") != "0"

WIDTH = 1024
TOPOLOGY = {
    # trunk blocks, branches: name -> (blocks, head)
    "lifter": (["res_common"], {"pose": (["res_pose1", "res_pose2", "res_pose3"], "downscale"),
                                "angle": (["res_angle1", "res_angle2", "res_angle3"], "angles")}),
    "predictor": ([], {"pose": (["res_pose1", "res_pose2", "res_pose3"], "downscale")}),
}


def _rup(x, m):
    return (x + m - 1) // m * m


class _Layer:
    __slots__ = ("name", "K", "N", "Kp", "Np", "W", "b", "gW", "gb", "Wb", "fused_ok", "off_W", "off_b")


class _Net:
    """Per-network parameter views + bf16 shadows."""

    def __init__(self, layers):
        self.layers = layers  # dict name -> _Layer


class MlpSet:
    def __init__(self, kind, in_dims, head_dims, max_rows, n_passes=1, device="cuda", train=True,
                 pass_branches=None, max_buckets=None, share_from=None, head_groups=None, zero_group=None):
        """in_dims[s]: input width of net s; head_dims[s]: dict head name -> width.
        max_rows: rows per pass; n_passes: forward passes per step sharing weights (1 or 2).
        pass_branches[p]: branches evaluated in pass p (default: all).
        head_groups: lists of (net, head) whose fp32 outputs share ONE [M, 32] buffer side by side (column offsets in list
        order): the geometry kernels read all depth / angle heads of a row from a single 128-byte line instead of one line
        per head (they index rows with the common leading dimension HEAD_LD, so a column-offset pointer is all they need).
        zero_group: a process group (one rank per GPU of ONE node) -> data parallelism WITHOUT gradient all-reduce of the big
        layers: rank r owns rows [r * 1024 / W, (r + 1) * 1024 / W) of every 1024 x 1024 weight matrix; the weight-gradient
        GEMM epilogues store their tiles as bf16 straight into the owner's staging buffer over NVLink (symmetric memory),
        the owner sums the W slots, runs Adam on its rows only and stores the new bf16 weights into every rank's shadow
        (push_*/zero_* methods; ZeRO-1 with both collectives fused into the producing kernels).
        share_from: another MlpSet of the same topology whose parameter / gradient / shadow buffers this set aliases
        (only the activation and gradient workspaces are private) -- several forward passes of one module can then be
        alive at once, each keeping its own activations for its own backward (utils/models_def.py)."""
        self.kind = kind
        self.trunk, self.branches = TOPOLOGY[kind]
        self.S = len(in_dims)
        self.M = max_rows
        self.n_passes = n_passes
        self.device = torch.device(device)
        self.train = train
        self.pass_branches = pass_branches or [list(self.branches) for _ in range(n_passes)]
        self.lib = _cabi.lib()
        dev = self.device
        # ---- parameter layout
        names = [("upscale", None)]
        for blk in self.trunk:
            names += [(blk + ".l1", None), (blk + ".l2", None)]
        for br, (blocks, head) in self.branches.items():
            for blk in blocks:
                names += [(blk + ".l1", None), (blk + ".l2", None)]
            names.append((head, br))
        self.layer_names = [n for n, _ in names]
        dims = {}
        for s in range(self.S):
            for n, br in names:
                if n == "upscale":
                    K, N = in_dims[s], WIDTH
                elif br is not None:
                    K, N = WIDTH, head_dims[s][n]
                else:
                    K, N = WIDTH, WIDTH
                dims[(s, n)] = (K, N)
        # Gradient buckets in the order the backward pass completes them (heads + deepest blocks first, upscale last).
        # The flat buffers are laid out bucket by bucket (all networks of a bucket adjacent), so a finished bucket is
        # ONE contiguous range: one all-reduce + one Adam + one shadow-cast launch, issued while backward continues.
        depth = max(len(blocks) for blocks, _ in self.branches.values())
        buckets = []
        for i in range(depth - 1, -1, -1):
            b = []
            for br, (blocks, head) in self.branches.items():
                if i == len(blocks) - 1:
                    b.append(head)
                if i < len(blocks):
                    b += [blocks[i] + ".l2", blocks[i] + ".l1"]
            buckets.append(b)
        tail = []
        for blk in reversed(self.trunk):
            tail += [blk + ".l2", blk + ".l1"]
        tail.append("upscale")
        buckets.append(tail)
        assert sorted(n for b in buckets for n in b) == sorted(self.layer_names)
        # fewer, larger buckets for latency-bound collectives (8 ranks: ~50-100 us fixed cost per all-reduce): merge
        # neighbours in completion order until at most max_buckets remain
        levels = [[i] for i in range(len(buckets))]
        self._n_levels = len(buckets)
        while max_buckets is not None and len(buckets) > max(1, max_buckets):
            sizes = [len(b) for b in buckets]
            i = min(range(len(buckets) - 1), key=lambda k: sizes[k] + sizes[k + 1])
            buckets[i:i + 2] = [buckets[i] + buckets[i + 1]]
            levels[i:i + 2] = [levels[i] + levels[i + 1]]
        self._bucket_levels = levels
        self.buckets = buckets
        # every weight / bias view starts on a 256-byte boundary of the flat buffers: the GEMM epilogue only takes
        # its vector path for 16-byte aligned operands (a 7x1024+7 head would otherwise misalign everything after it)
        order = [(s, n) for b in buckets for s in range(self.S) for n in b]
        total = sum(_rup(K * N, 64) + _rup(N, 64) for K, N in (dims[k] for k in order))
        self.n_params = total      # padded length of the flat buffers (padding stays zero through Adam)
        self.n_params_real = sum(K * N + N for K, N in dims.values())
        if share_from is not None:
            o = share_from
            assert (o.kind, o.S, o.n_params, o.buckets) == (kind, self.S, total, buckets) and (o.train or not train)
            self.master, self.nets, self.bucket_ranges, self.step_dev = o.master, o.nets, o.bucket_ranges, o.step_dev
            if train:
                self.grad, self.exp_avg, self.exp_avg_sq = o.grad, o.exp_avg, o.exp_avg_sq
        else:
            self.master = torch.zeros(total, dtype=torch.float32, device=dev)
            if train:
                self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
                self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
                self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        layers = [dict() for _ in range(self.S)]
        if share_from is None:
            self.bucket_ranges = []
            self.bucket_mid = []           # [start, mid): the big (1024 x 1024) weights; [mid, end): biases + small layers
        off = 0

        def fusable(K, N):
            """Weights whose gradient tiles are full 64-column slabs: their Adam step can run in the wgrad epilogue."""
            return K % 64 == 0 and N % 64 == 0

        for b in (buckets if share_from is None else []):
            start = off
            # Inside a bucket: first every big weight matrix (the wgrad GEMM may apply Adam to them in its epilogue and
            # then nothing else touches that range), then all biases and the small layers (one contiguous "rest" range
            # for the plain Adam kernel).  Every view starts on a 256-byte boundary.
            for s in range(self.S):
                for n in b:
                    K, N = dims[(s, n)]
                    L = _Layer()
                    L.name, L.K, L.N = n, K, N
                    L.Kp = _rup(K, 64)
                    L.Np = _rup(N, 64)
                    L.fused_ok = fusable(K, N)
                    L.Wb = torch.zeros(N, L.Kp, dtype=torch.bfloat16, device=dev)
                    layers[s][n] = L
                    if L.fused_ok:
                        L.off_W = off
                        off += _rup(N * K, 64)
            mid = off
            for s in range(self.S):
                for n in b:
                    L = layers[s][n]
                    if not L.fused_ok:
                        L.off_W = off
                        off += _rup(L.N * L.K, 64)
                    L.off_b = off
                    off += _rup(L.N, 64)
            for s in range(self.S):
                for n in b:
                    L = layers[s][n]
                    L.W = self.master[L.off_W:L.off_W + L.N * L.K].view(L.N, L.K)
                    L.gW = self.grad[L.off_W:L.off_W + L.N * L.K].view(L.N, L.K) if train else None
                    L.b = self.master[L.off_b:L.off_b + L.N]
                    L.gb = self.grad[L.off_b:L.off_b + L.N] if train else None
            self.bucket_ranges.append((start, off))
            self.bucket_mid.append(mid)
        if share_from is not None:
            self.bucket_mid = share_from.bucket_mid
        if share_from is None:
            self.nets = [_Net({n: layers[s][n] for n in self.layer_names}) for s in range(self.S)]
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # completed Adam steps (device counter)
        # ---- activation / gradient workspaces
        M = self.M
        P_ = n_passes
        bf = dict(dtype=torch.bfloat16, device=dev)

        def per_pass(t):
            return [t[p * M:(p + 1) * M] for p in range(P_)]
        self._x0buf = [torch.zeros(P_ * M, 64, **bf) for _ in range(self.S)]
        self.x0 = [[per_pass(self._x0buf[s])[p] for s in range(self.S)] for p in range(P_)]
        self._actbuf = [dict() for _ in range(self.S)]   # [s][name] -> [P*M, 1024]
        self.act = [[dict() for _ in range(self.S)] for _ in range(P_)]   # act[p][s][name] -> rows of pass p
        self.sign = []   # sign[p][s][blk] -> int32 [M, 32]
        self.head_out = []  # head_out[p][s][head] -> fp32 [M, HEAD_LD]
        all_blocks = list(self.trunk) + [b for blocks, _ in self.branches.values() for b in blocks]
        act_names = ["h0"] + [blk + sfx for blk in all_blocks for sfx in (".a1", ".y")]
        for s in range(self.S):
            for name in act_names:
                buf = torch.empty(P_ * M, WIDTH, **bf)
                self._actbuf[s][name] = buf
                for p in range(P_):
                    self.act[p][s][name] = buf[p * M:(p + 1) * M]
        for p in range(P_):
            sign_p, head_p = [], []
            for s in range(self.S):
                sign_p.append({blk: torch.zeros(M, WIDTH // 32, dtype=torch.int32, device=dev) for blk in all_blocks}
                              if train else {})
                head_p.append({head: torch.zeros(M, HEAD_LD, dtype=torch.float32, device=dev)
                               for _, head in self.branches.values()})
            for grp in (head_groups or []):
                width = sum(head_dims[s][h] for s, h in grp)
                assert width <= HEAD_LD, "packed heads do not fit one HEAD_LD-wide row"
                buf, col = torch.zeros(M, HEAD_LD, dtype=torch.float32, device=dev), 0
                for s, h in grp:
                    head_p[s][h] = buf[:, col:]           # row stride stays HEAD_LD
                    col += head_dims[s][h]
            self.sign.append(sign_p)
            self.head_out.append(head_p)
        if train:
            # gradients: G[p][s][layer] = rows of pass p of a row-major bf16 [P*M, N] buffer (heads: [P*M, 64]); dt per block
            self._Gbuf = [dict() for _ in range(self.S)]
            self.G = [[dict() for _ in range(self.S)] for _ in range(P_)]
            self.dt = [[dict() for _ in range(self.S)] for _ in range(P_)]
            self.E = [[torch.empty(M, WIDTH, **bf) for _ in range(self.S)] for _ in range(P_)]
            self.din = [[torch.zeros(M, HEAD_LD, dtype=torch.float32, device=dev) for _ in range(self.S)]
                        for _ in range(P_)]   # d/d(input part), fp32 [M, HEAD_LD]
            heads = [h for _, h in self.branches.values()]
            for s in range(self.S):
                for n in self.layer_names:
                    buf = torch.zeros(P_ * M, 64 if n in heads else WIDTH, **bf)
                    self._Gbuf[s][n] = buf
                    for p in range(P_):
                        self.G[p][s][n] = buf[p * M:(p + 1) * M]
                for p in range(P_):
                    for blk in all_blocks:
                        self.dt[p][s][blk] = torch.empty(M, WIDTH, **bf)
        self._plans = {}
        # learning rate of adam_step, read on the device (a captured graph follows set_lr without re-capture)
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._lr_host = None
        self.adam_hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self.zero = None
        if zero_group is not None:
            assert train and share_from is None
            self._enable_zero(zero_group)
        # B operand of the bias-gradient GEMMs: db = G^T . 1 (column 0 of a [rows, 16] matrix of ones)
        self._ones = torch.zeros(P_ * M, 16, **bf)
        self._ones[:, 0] = 1.0

    # ------------------------------------------------------------------------------------------
    # data parallelism by peer stores (reduce-scatter in the wgrad epilogue, sharded Adam, all-gather of the shadows)
    # ------------------------------------------------------------------------------------------
    def _enable_zero(self, pg):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        W, rank = dist.get_world_size(pg), dist.get_rank(pg)
        big = [(s, n) for b in self.buckets for s in range(self.S) for n in b if self.nets[s].layers[n].fused_ok]
        rows = cols = WIDTH
        assert all(self.nets[s].layers[n].N == rows and self.nets[s].layers[n].K == cols for s, n in big)
        assert 1 < W <= 8 and rows % (128 * W) == 0, "rows per owner must be a multiple of the 128-row CTA tile"
        rpo = rows // W
        owned = len(big) * rpo * cols                                   # elements of one staging slot
        dev = self.device
        stage = symm.empty(W * owned, dtype=torch.bfloat16, device=dev)
        stage.zero_()
        h_stage = symm.rendezvous(stage, pg)
        shadow = symm.empty(len(big) * rows * cols, dtype=torch.bfloat16, device=dev)
        shadow.zero_()
        h_shadow = symm.rendezvous(shadow, pg)
        table = (_cabi.AdamZeroLayer * len(big))()
        for li, (s, n) in enumerate(big):
            L = self.nets[s].layers[n]
            L.Wb = shadow[li * rows * cols:(li + 1) * rows * cols].view(rows, cols)      # shadows of the big layers: symmetric
            table[li].master_off = L.off_W
            table[li].stage_off = li * rpo * cols
            for r in range(W):
                table[li].shadow[r] = h_shadow.buffer_ptrs[r] + li * rows * cols * 2
        flags = symm.empty(64, dtype=torch.int32, device=dev)
        flags.zero_()
        h_flags = symm.rendezvous(flags, pg)
        flag_ptrs = (C.c_void_p * W)(*[h_flags.buffer_ptrs[r] for r in range(W)])
        tbytes = bytes(table)
        tdev = torch.frombuffer(bytearray(tbytes), dtype=torch.uint8).to(dev)
        self.zero = dict(pg=pg, W=W, rank=rank, big=big, index={k: i for i, k in enumerate(big)}, rpo=rpo, rows=rows, cols=cols,
                         owned=owned, stage=stage, h_stage=h_stage, shadow=shadow, h_shadow=h_shadow, table=tdev, n=len(big),
                         flags=flags, h_flags=h_flags, flag_ptrs=flag_ptrs)
        torch.cuda.synchronize()
        dist.barrier(group=pg)                 # every rank has zeroed its buffers before anybody stores into them

    def zero_barrier(self):
        """Device-side barrier over the ranks (symmetric-memory signal pads): everything the ranks stored into each
        other's buffers before it is visible after it."""
        z = self.zero
        check(self.lib.links_peer_barrier(z["flag_ptrs"], z["W"], z["rank"], 0, torch.cuda.current_stream().cuda_stream),
              "links_peer_barrier")

    def zero_adam(self):
        """Sharded optimiser step of the big layers: sum of the W staging slots -> Adam on the owned rows -> new bf16
        weights stored into every rank's shadow.  adam_prepare(grad_scale = 1 / W) must have run this step."""
        z = self.zero
        check(self.lib.links_adam_zero(self.master.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                       z["stage"].data_ptr(), z["owned"], z["table"].data_ptr(), z["n"], z["rpo"], z["cols"],
                                       z["W"], z["rank"], self.adam_hyper.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "links_adam_zero")

    def zero_sync_master(self):
        """Collective: make the fp32 master copies of the big layers complete on every rank (each rank only keeps its own
        rows up to date during training) -- call on ALL ranks before reading state_dict() for a checkpoint / validation."""
        import torch.distributed as dist
        z = self.zero
        if z is None:
            return
        W, rank, rpo, cols = z["W"], z["rank"], z["rpo"], z["cols"]
        mine = torch.cat([self.nets[s].layers[n].W[rank * rpo:(rank + 1) * rpo].reshape(-1) for s, n in z["big"]])
        full = torch.empty(W, mine.numel(), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(full, mine, group=z["pg"])
        blk = rpo * cols
        for li, (s, n) in enumerate(z["big"]):
            Wt = self.nets[s].layers[n].W
            for r in range(W):
                if r != rank:
                    Wt[r * rpo:(r + 1) * rpo].copy_(full[r, li * blk:(li + 1) * blk].view(rpo, cols))

    # ------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------
    def load_state_dicts(self, dicts):
        """dicts[s]: reference-style state dict (keys like 'res_pose1.l1.weight'); unused keys ignored."""
        with torch.no_grad():
            for s, sd in enumerate(dicts):
                for n, L in self.nets[s].layers.items():
                    L.W.copy_(sd[n + ".weight"].to(self.device, torch.float32))
                    L.b.copy_(sd[n + ".bias"].to(self.device, torch.float32))
        self.refresh_shadows()

    def state_dict(self, s):
        out = {}
        for n, L in self.nets[s].layers.items():
            out[n + ".weight"] = L.W.detach().clone()
            out[n + ".bias"] = L.b.detach().clone()
        return out

    def _cast_plan(self, bucket=None, rest_only=False):
        key = ("cast", bucket, rest_only)
        if key not in self._plans:
            names = self.layer_names if bucket is None else self.buckets[bucket]
            items = [(net.layers[n].W.data_ptr(), net.layers[n].Wb.data_ptr(), net.layers[n].N, net.layers[n].K,
                      net.layers[n].Kp) for net in self.nets for n in names
                     if not (rest_only and net.layers[n].fused_ok)]
            batches = []
            for i in range(0, len(items), _cabi.MAX_CAST_ITEMS):
                chunk = items[i:i + _cabi.MAX_CAST_ITEMS]
                arr = (_cabi.CastItem * len(chunk))()
                for j, (w, wb, N, K, ldw) in enumerate(chunk):
                    arr[j].W, arr[j].Wb, arr[j].N, arr[j].K, arr[j].ldw = w, wb, N, K, ldw
                batches.append((arr, len(chunk)))
            self._plans[key] = batches
        return self._plans[key]

    def refresh_shadows(self, bucket=None, rest_only=False):
        """fp32 master weights -> bf16 shadows (all layers, or the layers of one gradient bucket), one launch.
        rest_only: skip the big layers (the fused optimiser already refreshed their shadows)."""
        st = torch.cuda.current_stream().cuda_stream
        for arr, n in self._cast_plan(bucket, rest_only):
            check(self.lib.links_cast_weight_batched(arr, n, st), "links_cast_weight_batched")

    def compress_grads(self, bucket):
        """fp32 gradients of one bucket -> the bf16 communication buffer; returns the bf16 view to all-reduce."""
        if getattr(self, "grad16", None) is None:
            self.grad16 = torch.zeros(self.n_params, dtype=torch.bfloat16, device=self.device)
        a, b = self.bucket_ranges[bucket]
        st = torch.cuda.current_stream().cuda_stream
        check(self.lib.links_grad_compress_bf16(self.grad.data_ptr() + 4 * a, self.grad16.data_ptr() + 2 * a, b - a, st),
              "links_grad_compress_bf16")
        return self.grad16[a:b]

    def set_lr(self, lr):
        """Learning rate of the following adam_step calls, kept in a device word: captured graphs see new values."""
        if lr != self._lr_host:
            self.lr_dev.fill_(float(lr))
            self._lr_host = lr

    def adam_prepare(self, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, grad_scale=1.0):
        """Step constants of the fused optimiser (read by the wgrad epilogues of this step) -> self.adam_hyper."""
        self.set_lr(lr)
        check(self.lib.links_adam_prepare(self.step_dev.data_ptr(), self.lr_dev.data_ptr(), lr, betas[0], betas[1], eps,
                                          weight_decay, grad_scale, self.adam_hyper.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "links_adam_prepare")

    def adam_step(self, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, grad_scale=1.0, bucket=None,
                  last=True, grads_bf16=False, rest_only=False):
        """Adam on the whole flat buffer, or on one bucket's contiguous range.  The device-side step counter is
        advanced by the call with last=True (the other buckets of the same step pass last=False).
        rest_only (needs a bucket): only the biases and small layers -- the big weights of the bucket were updated by the
        fused optimiser in the wgrad epilogue (backward_ops(..., fuse_adam=True))."""
        self.set_lr(lr)               # no-op unless the value changed (never inside a captured step: see set_lr callers)
        st = torch.cuda.current_stream().cuda_stream
        a, b = (0, self.n_params) if bucket is None else self.bucket_ranges[bucket]
        if rest_only:
            a = self.bucket_mid[bucket]
        es = 4
        fn = self.lib.links_adam_step_g16 if grads_bf16 else self.lib.links_adam_step
        gptr = self.grad16.data_ptr() + 2 * a if grads_bf16 else self.grad.data_ptr() + a * es
        check(fn(self.master.data_ptr() + a * es, gptr, self.exp_avg.data_ptr() + a * es,
                 self.exp_avg_sq.data_ptr() + a * es, b - a, lr, betas[0], betas[1], eps, weight_decay,
                 0 if last else -1, self.step_dev.data_ptr(), grad_scale,
                 self.lr_dev.data_ptr(), st), "links_adam_step")
        self.refresh_shadows(bucket, rest_only)

    # ------------------------------------------------------------------------------------------
    # launch planning
    # ------------------------------------------------------------------------------------------
    def _launch(self, problems):
        """Returns a callable running one grouped launch (<= 8 problems each, split if more).  While a chain is being
        collected (_collect is a list) the problems become one LEVEL of the chain instead."""
        if getattr(self, "_collect", None) is not None:
            self._collect.append(list(problems))
            return ("level", len(self._collect) - 1)
        chunks = [problems[i:i + _cabi.MAX_GEMM_PROBLEMS] for i in range(0, len(problems), _cabi.MAX_GEMM_PROBLEMS)]
        arrs = [((GemmProblem * len(c))(*c), len(c)) for c in chunks]
        fn = self.lib.links_gemm_grouped

        def run():
            st = torch.cuda.current_stream().cuda_stream
            for arr, n in arrs:
                rc = fn(arr, n, st)
                if rc:
                    check(rc, "links_gemm_grouped")
        return run

    @staticmethod
    def _prob(A, B, M, N, K, lda, ldb, flags=0, bias=None, add0=None, add1=None, ymask=None, bits=None,
              sign_out=None, mid=None, out=None, out_f32=None, ld_f32=0, adam=None, push=None):
        P = GemmProblem()
        P.A, P.B, P.M, P.N, P.K, P.lda, P.ldb, P.flags = A.data_ptr(), B.data_ptr(), M, N, K, lda, ldb, flags
        P.bias = bias.data_ptr() if bias is not None else None
        for name, t in (("add0", add0), ("add1", add1), ("ymask", ymask), ("mid", mid), ("out", out)):
            if t is not None:
                setattr(P, name, t.data_ptr())
                setattr(P, "ld_" + name, t.stride(0))
        if bits is not None:
            P.bits, P.ld_bits = bits.data_ptr(), bits.stride(0)
        if sign_out is not None:
            P.sign_out, P.ld_sign = sign_out.data_ptr(), sign_out.stride(0)
        if out_f32 is not None:
            P.out_f32, P.ld_f32 = out_f32.data_ptr(), ld_f32 or out_f32.stride(0)
        if adam is not None:       # (p, m, v, shadow, hyper): fused optimiser instead of a stored gradient
            p_, m_, v_, sh, hy = adam
            P.adam_p, P.adam_m, P.adam_v, P.ld_f32 = p_.data_ptr(), m_.data_ptr(), v_.data_ptr(), p_.stride(0)
            P.adam_shadow, P.ld_shadow, P.adam_hyper = sh.data_ptr(), sh.stride(0), hy.data_ptr()
        if push is not None:       # (peer pointers, rows per owner, ld): reduce-scatter by peer stores
            ptrs, rpo, ld = push
            for r, q in enumerate(ptrs):
                P.push[r] = q
            P.push_rows, P.ld_push = rpo, ld
        return P

    def forward_plan(self, p, rows=None):
        """Launch list for pass p; inputs are self.x0[p][s] (bf16 [M,64]); outputs self.head_out[p][s]."""
        key = ("fwd", p, rows)
        if key not in self._plans:
            self._plans[key] = self._build_forward(p, rows)
        return self._plans[key]

    def _build_forward(self, p, rows=None):
        M = rows or self.M
        tr = self.train
        ops = []
        act = self.act[p]
        nets = self.nets
        # upscale (no activation, models_def.py:136)
        ops.append(self._launch([self._prob(self.x0[p][s], nets[s].layers["upscale"].Wb, M, WIDTH, 64, 64, 64,
                                            bias=nets[s].layers["upscale"].b, out=act[s]["h0"])
                                 for s in range(self.S)]))

        def block_ops(items):
            """items: list of (s, blk, xin_name)."""
            l1 = [self._prob(act[s][xin], nets[s].layers[blk + ".l1"].Wb, M, WIDTH, WIDTH, WIDTH, WIDTH,
                             flags=EPI_LEAKY_PRE, bias=nets[s].layers[blk + ".l1"].b, out=act[s][blk + ".a1"])
                  for s, blk, xin in items]
            l2 = [self._prob(act[s][blk + ".a1"], nets[s].layers[blk + ".l2"].Wb, M, WIDTH, WIDTH, WIDTH, WIDTH,
                             flags=EPI_LEAKY_PRE | EPI_LEAKY_POST, bias=nets[s].layers[blk + ".l2"].b,
                             sign_out=self.sign[p][s][blk] if tr else None, add0=act[s][xin],
                             out=act[s][blk + ".y"]) for s, blk, xin in items]
            ops.append(self._launch(l1))
            ops.append(self._launch(l2))

        x = "h0"
        for blk in self.trunk:
            block_ops([(s, blk, x) for s in range(self.S)])
            x = blk + ".y"
        active = self.pass_branches[p]
        depth = max(len(self.branches[br][0]) for br in active)
        xin = {br: x for br in active}
        for i in range(depth):
            items = []
            for br in active:
                blocks = self.branches[br][0]
                if i < len(blocks):
                    items += [(s, blocks[i], xin[br]) for s in range(self.S)]
            block_ops(items)
            for br in active:
                if i < len(self.branches[br][0]):
                    xin[br] = self.branches[br][0][i] + ".y"
        heads = []
        for br in active:
            head = self.branches[br][1]
            for s in range(self.S):
                L = nets[s].layers[head]
                heads.append(self._prob(act[s][xin[br]], L.Wb, M, L.N, WIDTH, WIDTH, WIDTH, bias=L.b,
                                        out_f32=self.head_out[p][s][head]))
        ops.append(self._launch(heads))
        return ops

    def backward_plan(self, p, need_input_grad, rows=None, wgrad=False, fuse_adam=False, push=False):
        """dgrad chain of pass p.  Inputs: self.G[p][s][head] (bf16 [M,64], zero beyond the head width) filled by the
        loss kernels.  Outputs: G of every layer, optionally self.din[p][s] = d/d(input part) fp32.
        dX = G . W reads the forward shadow W [N, Kp] as an MN-major B operand."""
        """With wgrad=True (the LAST pass to run backward) the weight-gradient GEMMs of a bucket are issued as soon as
        the dgrad chain has completed its G buffers, each followed by a ("bucket", b) marker: run(plan, on_bucket)
        calls on_bucket(b) there so the step driver can start the bucket's all-reduce / Adam on another stream."""
        key = ("bwd", p, need_input_grad, rows, wgrad, fuse_adam, push)
        if key not in self._plans:
            self._plans[key] = self._build_backward(p, need_input_grad, rows, wgrad, fuse_adam, push)
        return self._plans[key]

    def _build_backward(self, p, need_input_grad, rows=None, wgrad=False, fuse_adam=False, push=False):
        M = rows or self.M
        ops = []

        done = set()

        def bucket_done(level):
            """`level` indexes the un-merged completion order (deepest branch level = 0 ... trunk + upscale = last); the
            bucket that contains it is issued once its LAST level has been reached."""
            if not wgrad:
                return
            done.add(level)
            for b, members in enumerate(self._bucket_levels):
                if b not in self._issued and all(l in done for l in members):
                    self._issued.add(b)
                    ops.extend(self._wgrad_ops(b, rows, fuse_adam, push))
                    ops.append(("bucket", b))
        self._issued = set()
        act, G, dt, sign, nets = self.act[p], self.G[p], self.dt[p], self.sign[p], self.nets
        active = self.pass_branches[p]

        def xin_of(br, i):
            blocks = self.branches[br][0]
            if i > 0:
                return blocks[i - 1]
            return self.trunk[-1] if self.trunk else None

        # heads: dy_last = G_head . W_head ; masks of the last block of the branch
        probs = []
        for br in active:
            blocks, head = self.branches[br]
            last = blocks[-1]
            for s in range(self.S):
                L = nets[s].layers[head]
                probs.append(self._prob(G[s][head], L.Wb, M, WIDTH, L.N, 64, L.Kp, flags=GEMM_B_MN,
                                        ymask=act[s][last + ".y"], mid=dt[s][last], bits=sign[s][last],
                                        out=G[s][last + ".l2"]))
        ops.append(self._launch(probs))

        def l2_dgrad(items):
            ops.append(self._launch([
                self._prob(G[s][blk + ".l2"], nets[s].layers[blk + ".l2"].Wb, M, WIDTH, WIDTH, WIDTH, WIDTH,
                           flags=GEMM_B_MN, ymask=act[s][blk + ".a1"], out=G[s][blk + ".l1"])
                for s, blk in items]))

        def l1_dgrad(items):
            """items: (s, blk, prev, mode) -- mode 'mask' (prev is a block), 'raw' (-> E), 'merge' (add E), 'up'."""
            probs = []
            for s, blk, prev, mode in items:
                L = nets[s].layers[blk + ".l1"]
                kw = dict(add0=dt[s][blk], flags=GEMM_B_MN)
                if mode == "raw":
                    kw.update(out=self.E[p][s])
                elif mode == "up":
                    kw.update(out=G[s]["upscale"])
                else:
                    if mode == "merge":
                        kw.update(add1=self.E[p][s])
                    kw.update(ymask=act[s][prev + ".y"], mid=dt[s][prev], bits=sign[s][prev], out=G[s][prev + ".l2"])
                probs.append(self._prob(G[s][blk + ".l1"], L.Wb, M, WIDTH, WIDTH, WIDTH, WIDTH, **kw))
            ops.append(self._launch(probs))

        depth = max(len(self.branches[br][0]) for br in active)
        full_depth = max(len(blocks) for blocks, _ in self.branches.values())
        assert not wgrad or depth == full_depth, "the wgrad pass must run every branch"
        for i in range(depth - 1, -1, -1):
            brs = [br for br in active if i < len(self.branches[br][0])]
            l2_dgrad([(s, self.branches[br][0][i]) for br in brs for s in range(self.S)])
            if i > 0:
                l1_dgrad([(s, self.branches[br][0][i], self.branches[br][0][i - 1], "mask") for br in brs
                          for s in range(self.S)])
            else:
                prev = self.trunk[-1] if self.trunk else None
                if prev is None:
                    assert len(brs) == 1
                    l1_dgrad([(s, self.branches[brs[0]][0][0], None, "up") for s in range(self.S)])
                elif len(brs) == 1:
                    l1_dgrad([(s, self.branches[brs[0]][0][0], prev, "mask") for s in range(self.S)])
                else:
                    assert len(brs) == 2
                    l1_dgrad([(s, self.branches[brs[0]][0][0], prev, "raw") for s in range(self.S)])
                    l1_dgrad([(s, self.branches[brs[1]][0][0], prev, "merge") for s in range(self.S)])
            # G of level i (and of the heads for the deepest level) is final and the level's weights have been read for
            # the last time in this step: its bucket may be reduced / updated while the chain continues
            bucket_done(full_depth - 1 - i)
        for t in range(len(self.trunk) - 1, -1, -1):
            blk = self.trunk[t]
            l2_dgrad([(s, blk) for s in range(self.S)])
            if t > 0:
                l1_dgrad([(s, blk, self.trunk[t - 1], "mask") for s in range(self.S)])
            else:
                l1_dgrad([(s, blk, None, "up") for s in range(self.S)])
        if need_input_grad:
            probs = []
            for s in range(self.S):
                L = nets[s].layers["upscale"]
                probs.append(self._prob(G[s]["upscale"], L.Wb, M, L.K, WIDTH, WIDTH, L.Kp, flags=GEMM_B_MN,
                                        out_f32=self.din[p][s]))
            ops.append(self._launch(probs))
        bucket_done(self._n_levels - 1)
        return ops

    def _layer_input(self, name):
        """Name of the activation feeding layer `name` (or 'x0')."""
        if name == "upscale":
            return "x0"
        for br, (blocks, head) in self.branches.items():
            if name == head:
                return blocks[-1] + ".y"
        blk, l = name.rsplit(".", 1)
        if l == "l2":
            return blk + ".a1"
        if blk in self.trunk:
            t = self.trunk.index(blk)
            return "h0" if t == 0 else self.trunk[t - 1] + ".y"
        for br, (blocks, head) in self.branches.items():
            if blk in blocks:
                i = blocks.index(blk)
                if i > 0:
                    return blocks[i - 1] + ".y"
                return self.trunk[-1] + ".y" if self.trunk else "h0"
        raise KeyError(name)

    def _layer_passes(self, name):
        """Passes in which layer `name` is evaluated."""
        if name == "upscale" or name.split(".")[0] in self.trunk:
            return list(range(self.n_passes))
        for br, (blocks, head) in self.branches.items():
            if name == head or name.split(".")[0] in blocks:
                return [p for p in range(self.n_passes) if br in self.pass_branches[p]]
        raise KeyError(name)

    def _wgrad_ops(self, bucket, rows=None, fuse_adam=False, push=False):
        """dW = G^T . X of one bucket's layers, contracted over the rows of every pass that used the layer (G and X read
        as MN-major operands straight from their row-major buffers); db = G^T . 1 as one more (N = 1) GEMM problem per
        layer, written straight into the flat gradient buffer.  fuse_adam: the big layers' problems apply the optimiser
        step in their epilogue instead of storing dW (adam_prepare must have run; adam_step(rest_only=True) finishes the
        bucket)."""
        M = rows or self.M
        assert rows is None or self.n_passes == 1, "partial rows only supported for single-pass sets"
        probs = []
        for s in range(self.S):
            for n in self.buckets[bucket]:
                L = self.nets[s].layers[n]
                passes = self._layer_passes(n)
                assert passes == list(range(len(passes))), "passes using a layer must be a prefix"
                Kc = (len(passes) - 1) * self.M + M
                xin = self._layer_input(n)
                X = self._x0buf[s] if xin == "x0" else self._actbuf[s][xin]
                Gb = self._Gbuf[s][n]
                if push and L.fused_ok:
                    z = self.zero
                    li = z["index"][(s, n)]
                    base = (z["rank"] * z["owned"] + li * z["rpo"] * z["cols"]) * 2
                    ptrs = [z["h_stage"].buffer_ptrs[r] + base for r in range(z["W"])]
                    probs.append(self._prob(Gb, X, L.N, L.K, Kc, Gb.stride(0), X.stride(0), flags=GEMM_A_MN | GEMM_B_MN,
                                            push=(ptrs, z["rpo"], z["cols"])))
                elif fuse_adam and L.fused_ok:
                    off = L.off_W
                    adam = (L.W, self.exp_avg[off:off + L.N * L.K].view(L.N, L.K),
                            self.exp_avg_sq[off:off + L.N * L.K].view(L.N, L.K), L.Wb, self.adam_hyper)
                    probs.append(self._prob(Gb, X, L.N, L.K, Kc, Gb.stride(0), X.stride(0), flags=GEMM_A_MN | GEMM_B_MN,
                                            adam=adam))
                else:
                    probs.append(self._prob(Gb, X, L.N, L.K, Kc, Gb.stride(0), X.stride(0), flags=GEMM_A_MN | GEMM_B_MN,
                                            out_f32=L.gW, ld_f32=L.K))
                probs.append(self._prob(Gb, self._ones, L.N, 1, Kc, Gb.stride(0), 16, flags=GEMM_A_MN | GEMM_B_MN,
                                        out_f32=L.gb.view(L.N, 1), ld_f32=1))
        return [self._launch(probs)]

    def wgrad_plan(self, rows=None):
        """All weight / bias gradients (every bucket), for callers that do not interleave them with backward."""
        key = ("wgrad", rows)
        if key not in self._plans:
            self._plans[key] = self._build_wgrad(rows)
        return self._plans[key]

    def _build_wgrad(self, rows=None):
        ops = []
        for b in range(len(self.buckets)):
            ops.extend(self._wgrad_ops(b, rows))
        return ops

    @staticmethod
    def run(ops, on_bucket=None):
        for op in ops:
            if isinstance(op, tuple):
                if on_bucket is not None:
                    on_bucket(op[1])
            else:
                op()

    # ------------------------------------------------------------------------------------------
    # chain launches: a whole pass as ONE persistent kernel with tile-level dependencies
    # ------------------------------------------------------------------------------------------
    def _chain_op(self, levels):
        """levels: list of problem lists in dependency order -> callable running one links_gemm_chain_run.  Producers
        are found by address: an operand (A, add0, add1) that overlaps the out / mid / out_f32 range of an earlier
        problem of the chain depends on it -- row block by row block, or on ALL of its rows when the operand is an
        MN-major A (weight gradients contract over the rows)."""
        flat = [(lv, P) for lv, probs in enumerate(levels) for P in probs]
        n = len(flat)
        assert 0 < n <= _cabi.MAX_CHAIN_PROBLEMS, n
        arr = (ChainProblem * n)()
        written = []           # (begin, end, problem index)

        def extent(ptr, rows, ld, esize):
            return (ptr, ptr + rows * ld * esize)

        def producer(ptr, rows, ld, esize):
            if not ptr:
                return -1
            b, e = extent(ptr, rows, ld, esize)
            hits = sorted({i for (wb, we, i) in written if wb < e and b < we})
            assert len(hits) <= 1, "operand written by several problems of one chain"
            return hits[0] if hits else -1

        b_reads = []           # (begin, end, problem index) of every B operand (weights)
        for i, (lv, P) in enumerate(flat):
            cp = arr[i]
            C.memmove(C.byref(cp.g), C.byref(P), C.sizeof(GemmProblem))
            cp.level = lv
            a_mn = bool(P.flags & GEMM_A_MN)
            cp.dep[0] = producer(P.A, P.K if a_mn else P.M, P.lda, 2)
            cp.dep[1] = producer(P.add0, P.M, P.ld_add0, 2)
            cp.dep[2] = producer(P.add1, P.M, P.ld_add1, 2)
            cp.dep_all_rows = 1 if (a_mn and cp.dep[0] >= 0) else 0
            if P.adam_shadow:
                # write-after-read: the fused optimiser overwrites the layer's bf16 shadow, which earlier problems of the
                # chain (the dgrad of that layer) read as their B operand -- wait for all of their tiles
                sb, se = extent(P.adam_shadow, P.M, P.ld_shadow, 2)
                readers = sorted({j for (rb, re, j) in b_reads if rb < se and sb < re})
                assert len(readers) <= 2 and not P.add0 and not P.add1, "too many readers of a shadow inside one chain"
                for slot, j in zip((1, 2), readers):
                    cp.dep[slot] = j
                    cp.dep_all_rows |= 1 << slot
            b_mn = bool(P.flags & GEMM_B_MN)
            b_reads.append(extent(P.B, P.K if b_mn else P.N, P.ldb, 2) + (i,))
            for ptr, ld, es in ((P.out, P.ld_out, 2), (P.mid, P.ld_mid, 2), (P.out_f32, P.ld_f32, 4)):
                if ptr:
                    written.append(extent(ptr, P.M, ld, es) + (i,))
        L = self.lib
        nbytes = L.links_gemm_chain_ws_bytes(arr, n)
        if nbytes == 0:
            raise _cabi.LinksError("links_gemm_chain_ws_bytes rejected the chain")
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        plan = ChainPlan()
        st = torch.cuda.current_stream().cuda_stream
        check(L.links_gemm_chain_build(arr, n, ws.data_ptr(), nbytes, C.byref(plan), st), "links_gemm_chain_build")
        self._chain_keep = getattr(self, "_chain_keep", []) + [(ws, plan, arr)]
        fn = L.links_gemm_chain_run
        ref = C.byref(plan)

        def run():
            rc = fn(ref, torch.cuda.current_stream().cuda_stream)
            if rc:
                check(rc, "links_gemm_chain_run")
        run.plan = plan
        return run

    def _chained(self, key, build, split_at_buckets=False, max_ctas=None):
        """Re-plan `build()` (which issues self._launch calls and returns an op list) as chain launches: consecutive GEMM
        levels become one chain; the other ops (bias column sums, bucket markers) follow the chain that feeds them.  With
        split_at_buckets a chain ends at every bucket marker, so the bucket's all-reduce / Adam can start while the next
        chain runs (data parallelism); otherwise the whole op list is a single chain."""
        if key in self._plans:
            return self._plans[key]
        self._collect = []
        try:
            ops = build()
        finally:
            levels_all, self._collect = self._collect, None
        out, cur, deferred = [], [], []
        # max_ctas: the chain's persistent grid leaves (SMs - max_ctas) SMs to kernels that run NEXT to it (flows, NCCL):
        # a persistent grid that does not fit next to them would wait for them with its dependent tiles stalled
        prev = self.lib.links_gemm_set_max_ctas(max_ctas) if max_ctas else None

        def flush():
            if cur:
                out.append(self._chain_op([levels_all[i] for i in cur]))
            out.extend(deferred)
            del cur[:], deferred[:]
        try:
            for op in ops:
                if isinstance(op, tuple) and op[0] == "level":
                    cur.append(op[1])
                elif isinstance(op, tuple):                       # ("bucket", b)
                    deferred.append(op)
                    if split_at_buckets:
                        flush()
                else:
                    deferred.append(op)
            flush()
        finally:
            if max_ctas:
                self.lib.links_gemm_set_max_ctas(prev)
        self._plans[key] = out
        return out

    def forward_ops(self, p, rows=None, max_ctas=None):
        """Launch list of forward pass p: one chain launch (default) or the layer-by-layer grouped launches."""
        if not USE_CHAIN:
            return self.forward_plan(p, rows)
        return self._chained(("cfwd", p, rows), lambda: self._build_forward(p, rows), max_ctas=max_ctas)

    def backward_ops(self, p, need_input_grad, rows=None, wgrad=False, split_at_buckets=False, max_ctas=None,
                     fuse_adam=False, push=False):
        """dgrad chain of pass p (+ weight / bias gradients with wgrad=True; + the optimiser step of the big layers
        inside the wgrad epilogues with fuse_adam=True -- single-GPU steps only, gradients are then never stored)."""
        if not USE_CHAIN:
            return self.backward_plan(p, need_input_grad, rows, wgrad, fuse_adam, push)
        return self._chained(("cbwd", p, need_input_grad, rows, wgrad, split_at_buckets, fuse_adam, push),
                             lambda: self._build_backward(p, need_input_grad, rows, wgrad, fuse_adam, push), split_at_buckets,
                             max_ctas)

    def wgrad_ops(self, rows=None):
        if not USE_CHAIN:
            return self.wgrad_plan(rows)
        return self._chained(("cwgrad", rows), lambda: self._build_wgrad(rows))
