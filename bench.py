#!/usr/bin/env python
"""bench.py -- LInKs self-supervised lifter training step on B200 (BASELINE.json configs[1]).

One "step" = one leg/torso lifter step + one left/right lifter step (lift, rotate, reproject, per-part flow NLL,
re-lift, backward, Adam) on the same batch of B = 1024 synthetic 17-joint poses per GPU (N = 2B rows after the
flow-sampling concat).  Weak scaling: every rank owns B poses; lifter gradients are all-reduced over NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see the task contract): value = poses/s with inputs resident in HBM (CUDA-graph
replay), e2e = same through the public step API with pinned-host inputs copied H2D and losses read back D2H every
step, roofline = the tcgen05 GEMM (dominant kernel) timed in isolation with CUDA events, cpu_baseline = the CPU
oracle port on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "links-3d-human-pose-estimation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("WANDB_MODE", "disabled")

METRIC = "lifter_train_step_poses_per_sec"
UNIT = "poses/s"
# algorithmic FLOPs per DataLoader pose (SURVEY 8d / BASELINE.md 5): necessary work only
FLOP_PER_POSE = {"lt": 0.560e9, "lr": 0.561e9}


def gemm_flops_per_pose(kind):
    """2 * MACs of the lifter GEMMs per DataLoader pose (2 rows/pose): fwd pass 1 (pose+angle), fwd pass 2 (pose),
    dgrad (+ pass-2 upscale dgrad) and wgrad of both."""
    nj = {"lt": (7, 10), "lr": (11, 11)}[kind]
    tot = 0
    for n in nj:
        k = 2 * n
        body = 1024 * 1024
        full = k * 1024 + 14 * body + 1024 * n + 1024          # pass 1
        pose = k * 1024 + 8 * body + 1024 * n                   # pass 2
        fwd = full + pose
        dgrad = (14 * body + 1024 * n + 1024) + (8 * body + 1024 * n + k * 1024)
        wgrad = full + pose
        tot += fwd + dgrad + wgrad
    return 2 * 2 * tot   # 2 rows per pose, 2 FLOP per MAC


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = float(s[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_weights():
    from links_b200.init import init_flow_params, init_lifter_params
    nets = {"lt": [init_lifter_params(7, 11), init_lifter_params(10, 12)],
            "lr": [init_lifter_params(11, 13), init_lifter_params(11, 14)]}
    flows = {"lt": [init_flow_params(14, 41), init_flow_params(20, 42)],
             "lr": [init_flow_params(22, 43), init_flow_params(22, 44)]}
    full = init_flow_params(34, 40)
    return nets, flows, full


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle port of the same step on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_steps(batch, steps, warmup, seed=0):
    import torch
    from links_b200.synth import synth_poses
    from oracle import steps as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets, flows, full = make_weights()
    pn = {k: [OS.params_require_grad(p) for p in v] for k, v in nets.items()}
    opts = {k: OS.make_adam(v) for k, v in pn.items()}
    x2d, _ = synth_poses(batch, seed=1234 + seed)
    x = torch.from_numpy(x2d)
    g = torch.Generator().manual_seed(seed)
    times = []
    for it in range(warmup + steps):
        noise = torch.randn(batch, 34, generator=g)
        eps_x, u_y = torch.randn(2 * batch, generator=g), torch.rand(2 * batch, generator=g)
        t0 = time.perf_counter()
        u = OS.sample_poses(x, full, noise)
        for kind, fn in (("lt", OS.lt_step), ("lr", OS.lr_step)):
            for o in opts[kind]:
                o.zero_grad()
            out = fn(u, pn[kind][0], pn[kind][1], flows[kind][0], flows[kind][1], eps_x, u_y)
            out["loss"].backward()
            for o in opts[kind]:
                o.step()
            _ = out["loss"].item()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    sec, cores = cpu_steps(batch, args.steps, args.warmup)
    val = batch / sec
    sample = "oracle port (PyTorch CPU fp32), %d-pose batches (LT+LR step each), %d timed steps" % (batch, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LT+LR lifter training step, B=%d poses/step (bounded CPU sample of the B=1024 "
                                   "config)" % batch, "batch_per_step": batch},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    # Libraries (NCCL prints its version banner) write to stdout; the contract is ONE JSON line there.  Point fd 1 at
    # stderr for the whole run and keep the real stdout for the final line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (line + "\n").encode())

    import torch
    import torch.distributed as dist
    from links_b200 import _cabi
    from links_b200.steps import LifterStep, StepGroup
    from links_b200.synth import synth_poses

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        # the all-reduce kernels overlap the backward GEMMs: keep them on a few SMs and keep the GEMM grid off those
        if args.nccl_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
            os.environ.setdefault("NCCL_MIN_CTAS", str(args.nccl_ctas))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
        if args.gemm_ctas > 0:
            _cabi.lib().links_gemm_set_max_ctas(args.gemm_ctas)
    B = args.batch
    nets, flows, full = make_weights()
    # one communicator per branch: the LT and LR all-reduces are independent and may run concurrently
    pgs = {"lt": pg, "lr": pg}
    if world > 1 and not args.shared_comm:
        pgs["lr"] = dist.new_group(ranks=list(range(world)), backend="nccl")
    cfg = {"grad_comm": args.grad_comm, "dp_buckets": args.dp_buckets}
    steps = {k: LifterStep(k, B, nets[k], flows[k], full, cfg=cfg, process_group=pgs[k]) for k in ("lt", "lr")}
    lt, lr = steps["lt"], steps["lr"]

    # ---- inputs: per-rank shard of the global batch, pinned on the host for the e2e arm
    x2d, _ = synth_poses(B, seed=1234 + rank)
    gen = torch.Generator().manual_seed(1000 + rank)
    host = {"x": torch.from_numpy(x2d).pin_memory(), "noise": torch.randn(B, 34, generator=gen).pin_memory(),
            "eps_x": torch.randn(2 * B, generator=gen).pin_memory(), "u_y": torch.rand(2 * B, generator=gen).pin_memory()}
    host_losses = torch.zeros(2, 8).pin_memory()
    h2d_bytes = sum(t.numel() * 4 for t in host.values())
    d2h_bytes = host_losses.numel() * 4

    def upload():
        for s in (lt, lr):
            s.x.copy_(host["x"], non_blocking=True); s.noise.copy_(host["noise"], non_blocking=True)
            s.eps_x.copy_(host["eps_x"], non_blocking=True); s.u_y.copy_(host["u_y"], non_blocking=True)

    def download():
        host_losses[0].copy_(lt.losses, non_blocking=True)
        host_losses[1].copy_(lr.losses, non_blocking=True)

    group = StepGroup([lt, lr])      # the two independent lifter steps run as parallel branches of one graph

    def one_step():
        if args.serial:
            lt.step()
            lr.step()
        else:
            group.step()

    def stage(msg):
        if args.verbose:
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    # ---- count launches of one eager step (also serves as warm-up / lazy init)
    stage("built steps")
    upload()
    counter = _cabi.install_launch_counter()
    one_step()
    launches_per_step = counter.stop()
    torch.cuda.synchronize()
    stage("first eager step done")

    # ---- capture the whole step in a CUDA graph (falls back to eager launches if capture is unavailable)
    side = torch.cuda.Stream()
    graph = None
    if not args.no_graph:
        try:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                one_step()
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    one_step()
            torch.cuda.current_stream().wait_stream(side)
            graph = g
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print("[bench] CUDA graph capture failed (%s); timing eager launches" % e, file=sys.stderr)
            graph = None
    run = (lambda: graph.replay()) if graph is not None else one_step
    stage("graph captured: %s" % (graph is not None))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage("warm-up done")
    ms_total = timed(run, K)
    stage("timed region done")
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end: pinned host inputs -> H2D, step, losses -> D2H, every step
    def e2e_step():
        upload()
        run()
        download()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    final_losses = {"lt": lt.loss_dict(), "lr": lr.loss_dict()}

    # ---- roofline of the dominant kernel (tcgen05 grouped GEMM): all GEMM launches of one step, timed alone
    def gemm_only():
        for s in (lt, lr):
            m = s.mlp
            for plan in (m.forward_plan(0), m.forward_plan(1), m.backward_plan(1, True), m.backward_plan(0, False)):
                m.run(plan)
            m.run(m.wgrad_plan()[0::2])          # the GEMM launches of every bucket (odd entries: bias column sums)
    counter = _cabi.install_launch_counter()
    gemm_only()
    n_gemm_launches = counter.stop()
    for _ in range(3):
        gemm_only()
    ms_gemm = timed(gemm_only, K) / K
    peaks = load_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch_mean")
    gemm_flops = (gemm_flops_per_pose("lt") + gemm_flops_per_pose("lr")) * B
    achieved = gemm_flops / (ms_gemm * 1e-3) / 1e12

    ms_step = ms_total / K
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / K * 1e-3)
    if rank == 0:
        cpu_sec, cores = cpu_steps(args.cpu_batch, 3, 1) if not args.skip_cpu else (None, os.cpu_count())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: leg/torso + left/right lifter self-supervised training step, "
                                   "B=%d poses per GPU (N=%d rows), 17 joints" % (B, 2 * B),
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "cuda_graph": graph is not None,
                       "branches": "LT and LR steps serial on one stream" if args.serial else
                                   "LT and LR steps as parallel branches (2 streams) of one CUDA graph",
                       "l2_policy": "no explicit flush: each step streams ~%.1f GB of activations/weights/gradients, "
                                    "far above the 126 MB L2" % (2 * 2048 * 1024 * 2 * 2 * 60 * (B / 1024) / 1e9),
                       "operands": "bf16 operands + bf16-stored activations, fp32 accumulate / master weights / losses",
                       "elevation_stats": "local shard" if world > 1 else "global",
                       "grad_allreduce": ("%s buckets over NCCL, overlapped with backward" % args.grad_comm) if world > 1 else "none",
                       "final_losses": final_losses},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
                         "traffic_note": "mean dram read+write bytes per GEMM launch from the ncu --set full capture in "
                                         "profiles/ (B=1024 config); algorithmic operand bytes of a 4-problem launch: 24 MB",
                         "flops_per_launch": gemm_flops / n_gemm_launches,
                         "us_per_launch": ms_gemm * 1e3 / n_gemm_launches,
                         "kernel": "links::gemm_grouped_kernel (tcgen05/TMEM/TMA), %d launches per step timed in "
                                   "isolation" % n_gemm_launches,
                         "ms_per_step_gemm_only": ms_gemm, "flops_per_step": gemm_flops,
                         "peak_source": "%s cuBLAS bf16 burst (kernel timed alone)" % peaks["source"]},
            "cpu_baseline": None if cpu_sec is None else {
                "value": args.cpu_batch / cpu_sec, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "oracle port (PyTorch CPU fp32), LT+LR step on %d-pose batches, 3 timed steps" % args.cpu_batch},
        }
        emit(json.dumps(line))
    if world > 1:
        # captured graphs hold NCCL work; drop them before the communicator goes away.  Tearing the process group down
        # with captured collectives alive was observed to hang on exit, so leave without running destructors.
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024, help="poses per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=256, help="poses per step of the bounded CPU sample")
    ap.add_argument("--impl", default="links_b200", choices=["links_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--dp-buckets", type=int, default=2, help="gradient buckets per network set under data parallelism")
    ap.add_argument("--shared-comm", action="store_true", help="one NCCL communicator for both branches")
    ap.add_argument("--nccl-ctas", type=int, default=0, help="CTAs (SMs) the NCCL all-reduce kernels may use (0: NCCL default)")
    ap.add_argument("--gemm-ctas", type=int, default=-1, help="GEMM grid cap under data parallelism (<= 0: no cap)")
    ap.add_argument("--grad-comm", default="bf16", choices=["bf16", "fp32"],
                    help="dtype of the data-parallel gradient all-reduce (bf16 = compressed buckets)")
    ap.add_argument("--serial", action="store_true", help="run the LT and LR steps back to back on one stream")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
