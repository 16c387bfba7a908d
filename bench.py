#!/usr/bin/env python
"""bench.py -- LInKs self-supervised lifter training step on B200 (BASELINE.json configs[1]).

One "step" = one leg/torso lifter step + one left/right lifter step (lift, rotate, reproject, per-part flow NLL,
re-lift, backward, Adam) on the same batch of B = 1024 synthetic 17-joint poses per GPU (N = 2B rows after the
flow-sampling concat).  Weak scaling: every rank owns B poses; lifter gradients are all-reduced over NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see the task contract):
  value        poses/s with inputs resident in HBM (CUDA-graph replay of the merged LT+LR step)
  e2e          same through the public step API with pinned-host inputs copied H2D and losses read back D2H every step
  roofline     the tcgen05 GEMM (dominant kernel): every GEMM launch of one step timed alone with CUDA events
  parity       first-step losses at THIS batch size vs the CPU oracle (tolerance 1e-3 relative)
  cpu_baseline the CPU oracle port of the same step at the SAME batch on this box's host cores
  configs      short driver-visible timings of BASELINE configs #1 (flow step B=256), #4 (occlusion step, 4096 poses
               global) and #5 (sharded eval, 1.25 M poses per GPU, H2D included), each with its own roofline fraction
  strong_scaling  the same lifter step at a FIXED global batch of 8192 (config #3 read as strong scaling)
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
# torchrun exports OMP_NUM_THREADS=1 to every rank.  Rank 0 alone runs the CPU legs (the oracle step behind `parity`, the whole
# `--impl reference` arm) and those want every host core: lift the cap for rank 0 before torch initialises its thread pools.
if (os.environ.get("RANK", "0") == "0" and os.environ.get("OMP_NUM_THREADS") == "1"
        and int(os.environ.get("WORLD_SIZE", "1")) > 1):
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

METRIC = "lifter_train_step_poses_per_sec"
UNIT = "poses/s"
BODY = 1024 * 1024


def gemm_flops_per_pose(kind):
    """2 * MACs of the lifter GEMMs per DataLoader pose (2 rows/pose): fwd pass 1 (pose+angle), fwd pass 2 (pose),
    dgrad (+ pass-2 upscale dgrad) and wgrad of both."""
    nj = {"lt": (7, 10), "lr": (11, 11)}[kind]
    tot = 0
    for n in nj:
        k = 2 * n
        full = k * 1024 + 14 * BODY + 1024 * n + 1024          # pass 1
        pose = k * 1024 + 8 * BODY + 1024 * n                   # pass 2
        fwd = full + pose
        dgrad = (14 * BODY + 1024 * n + 1024) + (8 * BODY + 1024 * n + k * 1024)
        wgrad = full + pose
        tot += fwd + dgrad + wgrad
    return 2 * 2 * tot   # 2 rows per pose, 2 FLOP per MAC


def lifter_pose_macs(n):
    """MACs per row of the pose branch only (eval / occlusion use the lifters without the angle branch)."""
    return 2 * n * 1024 + 8 * BODY + 1024 * n


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
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
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
    nets = [init_lifter_params(7, 11), init_lifter_params(10, 12), init_lifter_params(11, 13), init_lifter_params(11, 14)]
    flows = [init_flow_params(14, 41), init_flow_params(20, 42), init_flow_params(22, 43), init_flow_params(22, 44)]
    full = init_flow_params(34, 40)
    return nets, flows, full


def make_inputs(B, rank, n_batches=2):
    """Per-rank batches of synthetic poses + the step's random draws (host tensors)."""
    import torch
    from links_b200.synth import synth_poses
    out = []
    for i in range(n_batches):
        x2d, _ = synth_poses(B, seed=1234 + 97 * i + rank)
        gen = torch.Generator().manual_seed(1000 + 31 * i + rank)
        out.append({"x": torch.from_numpy(x2d), "noise": torch.randn(B, 34, generator=gen),
                    "eps_x": torch.randn(2 * B, generator=gen), "u_y": torch.rand(2 * B, generator=gen)})
    return out


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle port of the same step on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_steps(batch, steps, warmup, first_losses=False):
    """LT + LR step of the CPU oracle (PyTorch fp32, all host cores) on `batch` poses: seconds per step, core count and
    (optionally) the first step's losses on bench.py's own first batch (the `parity` check)."""
    import torch
    from oracle import steps as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets, flows, full = make_weights()
    pn = [OS.params_require_grad(p) for p in nets]
    opts = OS.make_adam(pn)
    data = make_inputs(batch, 0)
    times, first = [], None
    for it in range(warmup + steps):
        d = data[0] if it == 0 else data[1]
        t0 = time.perf_counter()
        u = OS.sample_poses(d["x"], full, d["noise"])
        losses = {}
        for kind, fn, s0 in (("lt", OS.lt_step, 0), ("lr", OS.lr_step, 2)):
            for o in opts[s0:s0 + 2]:
                o.zero_grad()
            out = fn(u, pn[s0], pn[s0 + 1], flows[s0], flows[s0 + 1], d["eps_x"], d["u_y"])
            out["loss"].backward()
            for o in opts[s0:s0 + 2]:
                o.step()
            losses[kind] = {k: v.item() for k, v in out.items()}
        dt = time.perf_counter() - t0
        if it == 0:
            first = losses
        if it >= warmup:
            times.append(dt)
    return sum(times) / max(len(times), 1), cores, (first if first_losses else None)


def cpu_flow_steps(batch, steps, warmup):
    """BASELINE config #1 on the host cores: the full-pose flow training step of the CPU oracle (train_full_pose_norm_flow.py:
    67-98: NLL of the data + NLL of the flow's own noisy samples, Adam) -> the `cpu_baseline` object of configs[0]."""
    import torch
    from links_b200 import init as INIT
    from links_b200.synth import synth_poses
    from oracle import steps as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = INIT.init_flow_params(34, 40)
    pn = OS.params_require_grad(params)
    for k in pn:
        if "w_perm" in k:
            pn[k].requires_grad_(False)
    opt = torch.optim.Adam([v for v in pn.values() if v.requires_grad], lr=2e-4, weight_decay=1e-5)
    x2d, _ = synth_poses(batch, seed=77)
    x = torch.from_numpy(x2d)
    gen = torch.Generator().manual_seed(5)
    times = []
    for it in range(warmup + steps):
        noise = torch.randn(batch, 34, generator=gen)
        t0 = time.perf_counter()
        opt.zero_grad()
        out = OS.flow_step(x, pn, noise)
        out["loss"].backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": batch / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "cores": cores, "kind": "port",
            "sample": "oracle port (PyTorch CPU fp32) of the flow training step, B=%d, %d timed steps after %d warm-ups"
                      % (batch, steps, warmup)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch
    sec, cores, _ = cpu_steps(batch, args.steps, args.warmup)
    val = batch / sec
    sample = "oracle port (PyTorch CPU fp32), %d-pose batches (LT+LR step each), %d timed steps" % (batch, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: leg/torso + left/right lifter self-supervised training step, "
                                   "B=%d poses per step (N=%d rows), 17 joints" % (batch, 2 * batch),
                       "batch_per_gpu": batch, "global_batch": batch, "parallelism": "cpu"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def __call__(self, fn, n):
        """n calls of fn bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks (ms)."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = t.item()
        return ms


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def extra_configs(args, timed, world, rank, pg, peaks):
    """Driver-visible numbers for BASELINE configs #1, #4, #5 (short runs; each with its own roofline fraction)."""
    import torch
    from links_b200 import init as INIT
    from links_b200.flowtrain import FlowTrainStep
    from links_b200.occlusion import OCC_IN, OCC_NAMES, OCC_OUT, EvalRunner, OcclusionStep
    from links_b200.synth import synth_poses
    out = []
    burst = peaks["bf16_burst"]
    # ---- config #1: full-pose GLOW flow NLL fwd/bwd + Adam, batch 256 (train_full_pose_norm_flow.py:67-98); replicas
    B1 = 256
    ft = FlowTrainStep(34, INIT.init_flow_params(34, 40), B1, lr=2e-4)
    x2d, _ = synth_poses(B1, seed=77 + rank)
    ft.x.copy_(torch.from_numpy(x2d)); ft.noise.normal_()
    for _ in range(4):
        ft.run()                              # eager, capture + replay, replays (what the drop-in script does)
    n1 = 20
    ms = timed(ft.run, n1) / n1
    # per data pose: sampling (fwd + rev = 2 subnet passes) on 1 row, NLL fwd + reversible bwd (3 passes) on 2 rows,
    # parameter-gradient GEMMs (hidden recompute K=64 padded, dgrad, 2 wgrads) on 2 rows
    macs_pass = 427040
    flops1 = 2.0 * (2 * macs_pass + 2 * 3 * macs_pass + 2 * 8 * (64 * 1024 + 3 * 34 * 1024))
    out.append({"config": "configs[0]: full-pose GLOW flow training step (sample, NLL fwd/bwd on [x ; s], Adam), B=256 per GPU, "
                          "independent replicas, CUDA-graph replay",
                "value": world * B1 / (ms * 1e-3), "unit": "poses/s", "ms_per_step": ms, "n_gpus": world,
                "roofline": {"bound": "tensor", "achieved": flops1 * B1 / (ms * 1e-3) / 1e12, "peak": burst, "unit": "TFLOP/s",
                             "frac": flops1 * B1 / (ms * 1e-3) / 1e12 / burst,
                             "note": "whole step; 512 rows = 4 row tiles of the flow kernel: latency-bound by construction "
                                     "(8 coupling blocks x 3 passes run back to back per tile)"}})
    del ft
    # ---- config #4: occlusion-model training step, 4096 poses global (train_occlusion_models.py:144-314)
    B4 = max(2, (4096 // world) // 2 * 2)
    lifters = [INIT.init_lifter_params(7, 11), INIT.init_lifter_params(10, 12)]
    preds = {n: INIT.init_predictor_params(OCC_IN[n] // 3, OCC_OUT[n], 100 + i) for i, n in enumerate(OCC_NAMES)}
    oc = OcclusionStep(B4, lifters, preds, cfg={"grad_comm": "bf16" if args.grad_comm == "push" else args.grad_comm,
                                                "dp_buckets": args.dp_buckets}, process_group=pg)
    x2d, _ = synth_poses(B4, seed=55 + rank)
    oc.x.copy_(torch.from_numpy(x2d)); oc.u_y[0].uniform_(); oc.u_y[1].uniform_()
    for _ in range(3):
        oc.step()
    run4 = oc.step
    if world == 1 and not args.no_graph:          # what the drop-in script does (harness.run_training): graph replay
        oc.capture(warmup=0)
        run4 = oc.graph.replay
        run4()
    n4 = 10
    ms = timed(run4, n4) / n4
    pred_macs = sum(OCC_IN[n] * 1024 + 6 * BODY + 1024 * OCC_OUT[n] for n in OCC_NAMES)
    flops4 = 2.0 * (lifter_pose_macs(7) + lifter_pose_macs(10) + 3 * 3 * pred_macs)      # 3 rounds x (fwd, dgrad, wgrad)
    out.append({"config": "configs[3]: occlusion-model training step (8 predictors, 3 rounds, Adam), global batch %d = %d per "
                          "GPU, %s" % (B4 * world, B4, "CUDA-graph replay" if run4 != oc.step else "bf16 gradient all-reduce"),
                "value": world * B4 / (ms * 1e-3), "unit": "poses/s", "ms_per_step": ms, "n_gpus": world,
                "roofline": {"bound": "tensor", "achieved": flops4 * B4 / (ms * 1e-3) / 1e12, "peak": burst, "unit": "TFLOP/s",
                             "frac": flops4 * B4 / (ms * 1e-3) / 1e12 / burst, "note": "whole step incl. Adam / all-reduce"}})
    del oc
    # ---- config #5: sharded eval, 1.25 M poses per GPU (10 M over 8), pinned host -> H2D -> lift -> N-MPJPE / PA-MPJPE
    n5, chunk = args.eval_poses, 65536
    x2d, gt = synth_poses(chunk, seed=99 + rank)
    reps = (n5 + chunk - 1) // chunk
    hx, hg = torch.from_numpy(x2d).pin_memory(), torch.from_numpy(gt).pin_memory()       # one pinned chunk, re-sent
    ev = EvalRunner("lr", [INIT.init_lifter_params(11, 13), INIT.init_lifter_params(11, 14)], chunk=chunk, process_group=pg)
    dx = [torch.empty(chunk, 34, device="cuda") for _ in range(2)]
    dg = [torch.empty(chunk, 51, device="cuda") for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    cp_done = [torch.cuda.Event(), torch.cuda.Event()]

    def eval_pass():
        ev.reset()
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)
        for i in range(reps):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(ev_done[b])
                dx[b].copy_(hx, non_blocking=True); dg[b].copy_(hg, non_blocking=True)
                cp_done[b].record(copy_stream)
            main.wait_event(cp_done[b])
            ev.run_chunk(dx[b], dg[b])
            ev_done[b].record(main)
    eval_pass()
    ms = timed(eval_pass, 1)
    res = ev.result()
    n_done = reps * chunk
    flops5 = 2.0 * 2 * lifter_pose_macs(11)
    out.append({"config": "configs[4]: batched eval (lift + N-MPJPE + PA-MPJPE), %d poses per GPU in %d-pose chunks from pinned "
                          "host memory (H2D inside the timed region, double-buffered), one final reduction" % (n_done, chunk),
                "value": world * n_done / (ms * 1e-3), "unit": "poses/s", "ms_total": ms, "n_gpus": world,
                "h2d_bytes": n_done * 85 * 4, "pa_mpjpe_mm": res["pa_mpjpe"], "n_mpjpe_mm": res["n_mpjpe"],
                "roofline": {"bound": "tensor", "achieved": flops5 * n_done / (ms * 1e-3) / 1e12, "peak": burst,
                             "unit": "TFLOP/s", "frac": flops5 * n_done / (ms * 1e-3) / 1e12 / burst,
                             "note": "lifter GEMMs dominate (33.7 MFLOP per pose); the fused lift+score kernel moves "
                                     "408 B per pose"}})
    del ev
    torch.cuda.empty_cache()
    return out


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
    from links_b200 import mlp as MLP
    from links_b200.steps import LifterStep

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
    timed = Timer(world)
    B = args.batch
    nets, flows, full = make_weights()
    cfg = {"grad_comm": args.grad_comm, "dp_buckets": args.dp_buckets, "prefetch_sample": not args.no_prefetch,
           "global_elevation_stats": args.global_elevation_stats,
           "store_rot_2d": False,      # the full projected poses are a debugging output; the step consumes the part gathers
           "nccl_ctas": args.nccl_ctas}

    def stage(msg):
        if args.verbose:
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    try:
        step = LifterStep("both", B, nets, flows, full, cfg=cfg, process_group=pg)
    except Exception as e:  # noqa: BLE001
        if not (world > 1 and cfg["grad_comm"] == "push"):
            raise
        # symmetric memory (peer mapping over NVLink) unavailable on this box: fall back to the bucketed NCCL all-reduce
        print("[bench rank %d] push mode unavailable (%s: %s); falling back to bf16 NCCL buckets" % (rank, type(e).__name__, e),
              file=sys.stderr, flush=True)
        args.grad_comm = cfg["grad_comm"] = "bf16"
        cfg["dp_buckets"] = args.dp_buckets
        step = LifterStep("both", B, nets, flows, full, cfg=cfg, process_group=pg)
    data = make_inputs(B, rank)
    host = [{k: v.pin_memory() for k, v in d.items()} for d in data]
    host_losses = torch.zeros(2, 8).pin_memory()
    h2d_bytes = sum(t.numel() * 4 for t in host[0].values())
    d2h_bytes = host_losses.numel() * 4

    def upload(i):
        """Sampling inputs of the NEXT step + rotation draws of THIS step (sampling runs one step ahead)."""
        nxt = host[1]
        cur = host[0] if i == 0 else host[1]
        src_s = nxt if step.prefetch else cur
        step.x.copy_(src_s["x"], non_blocking=True); step.noise.copy_(src_s["noise"], non_blocking=True)
        step.eps_x.copy_(cur["eps_x"], non_blocking=True); step.u_y.copy_(cur["u_y"], non_blocking=True)

    def download():
        for i, k in enumerate(step.K):
            host_losses[i].copy_(k.losses, non_blocking=True)

    # ---- first step from fresh weights, eager: parity sample + launch count (also builds every plan)
    stage("built step")
    if step.prefetch:
        step.x.copy_(host[0]["x"]); step.noise.copy_(host[0]["noise"])
        step.prime()
    upload(0)
    counter = _cabi.install_launch_counter()
    step.step()
    launches_per_step = counter.stop()
    gemm_launches_per_step = counter.gemm
    torch.cuda.synchronize()
    first_losses = step.loss_dict()
    stage("first eager step done")

    # ---- capture the whole step in a CUDA graph
    graph = None
    upload(1)
    if not args.no_graph:
        try:
            step.capture(warmup=1)
            graph = step.graph
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print("[bench] CUDA graph capture failed (%s); timing eager launches" % e, file=sys.stderr)
            graph = None
    run = (lambda: graph.replay()) if graph is not None else step.step
    stage("graph captured: %s" % (graph is not None))

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
        upload(1)
        run()
        download()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    final_losses = step.loss_dict()

    # ---- roofline of the dominant kernel (tcgen05 GEMM chain kernel), timed alone with CUDA events:
    #   (a) `roofline`: the step's four chain launches as pure GEMM work on the whole machine -- the unfused plans (the
    #       weight-gradient tiles store dW instead of running Adam's 26 B/parameter read-modify-write in their epilogue)
    #       with no SMs set aside for the flow kernels;
    #   (b) `roofline.in_step`: the very launches the captured step issues (SM reservation next to the flows, optimiser
    #       fused into the weight-gradient epilogues), same FLOP count.
    m = step.mlp
    split = world > 1
    step_plans = [m.forward_ops(0), m.forward_ops(1, max_ctas=step._ctas_window),
                  m.backward_ops(1, need_input_grad=True, max_ctas=step._ctas_window),
                  m.backward_ops(0, need_input_grad=False, wgrad=True, split_at_buckets=split, max_ctas=step._ctas_tail,
                                 fuse_adam=step._fuse_adam)]
    if MLP.USE_CHAIN:
        pure_plans = [m._chained(("roof", 0), lambda: m._build_forward(0)), m._chained(("roof", 1), lambda: m._build_forward(1)),
                      m._chained(("roof", 2), lambda: m._build_backward(1, True)),
                      m._chained(("roof", 3), lambda: m._build_backward(0, False, None, True, False))]
        sel = lambda plans: [op for plan in plans for op in plan if hasattr(op, "plan")]
        gemm_ops, step_ops = sel(pure_plans), sel(step_plans)
        sim = {"chain_sim_us": [float(op.plan.sim_units) * 0.44 for op in gemm_ops],
               "chain_ideal_us": [float(op.plan.ideal_units) * 0.44 for op in gemm_ops],
               "chain_tiles": [int(op.plan.total_tiles) for op in gemm_ops]}
    else:
        pure_plans = [m.forward_plan(0), m.forward_plan(1), m.backward_plan(1, True), m.backward_plan(0, False, None, True, False)]
        sel = lambda plans: [op for plan in plans for op in plan if not isinstance(op, tuple)]
        gemm_ops, step_ops, sim = sel(pure_plans), sel(step_plans), {}

    def time_ops(ops):
        def run_ops():
            for op in ops:
                op()
        counter = _cabi.install_launch_counter()
        run_ops()
        counter.stop()
        for _ in range(3):
            run_ops()
        return timed(run_ops, K) / K, counter.gemm
    ms_gemm, n_gemm_launches = time_ops(gemm_ops)
    ms_gemm_step, n_step_launches = time_ops(step_ops)
    peaks = load_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch_mean")
    gemm_flops = (gemm_flops_per_pose("lt") + gemm_flops_per_pose("lr")) * B
    achieved = gemm_flops / (ms_gemm * 1e-3) / 1e12
    in_step = {"ms_per_step_gemm_launches": ms_gemm_step, "launches_per_step": n_step_launches,
               "achieved": gemm_flops / (ms_gemm_step * 1e-3) / 1e12,
               "frac": gemm_flops / (ms_gemm_step * 1e-3) / 1e12 / peaks["bf16_burst"],
               "note": "the launches of the captured step, timed alone: forward pass 2 and its dgrad chain run on %s SMs (the "
                       "rest is left to the concurrent part-flow NLL kernels), the tail chain on %s SMs with Adam fused into "
                       "the weight-gradient epilogues (HBM-bound: +26 B per parameter)" % (step._ctas_window, step._ctas_tail)}

    ms_step = ms_total / K
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / K * 1e-3)

    # ---- the other BASELINE configs and the strong-scaling reading of config #3 (short runs)
    del graph
    step.graph = None
    extras, strong = [], None
    if not args.skip_extra:
        stage("extra configs")
        del step, m, step_plans, pure_plans, gemm_ops, step_ops
        torch.cuda.empty_cache()
        Bs = max(2, (8192 // world) // 2 * 2)
        st2 = LifterStep("both", Bs, nets, flows, full, cfg=cfg, process_group=pg)
        d2 = make_inputs(Bs, rank, n_batches=1)[0]
        st2.x.copy_(d2["x"]); st2.noise.copy_(d2["noise"]); st2.eps_x.copy_(d2["eps_x"]); st2.u_y.copy_(d2["u_y"])
        if st2.prefetch:
            st2.prime()
        st2.capture(warmup=2)
        for _ in range(3):
            st2.replay()
        ks = 10
        ms_s = timed(st2.replay, ks) / ks
        strong = {"global_batch": Bs * world, "batch_per_gpu": Bs, "n_gpus": world, "ms_per_step": ms_s,
                  "value": Bs * world / (ms_s * 1e-3), "unit": UNIT,
                  "gemm_tflops_in_step": (gemm_flops_per_pose("lt") + gemm_flops_per_pose("lr")) * Bs / (ms_s * 1e-3) / 1e12}
        st2.graph = None
        del st2
        torch.cuda.empty_cache()
        extras = extra_configs(args, timed, world, rank, pg, peaks)

    if rank == 0:
        parity, cpu_base = None, None
        if not args.skip_cpu:
            # the timed CPU leg belongs to rank 0 at N = 1 only (~10 s of CPU work); data-parallel runs keep the single
            # oracle step the `parity` object needs (measured at N = 2 / 8: the same leg beside the other ranks' spinning
            # host threads is 20-30 x slower and would add minutes to every scaling run)
            n_cpu = 15 if world == 1 else 0
            cpu_sec, cores, ref_first = cpu_steps(B, n_cpu, 1, first_losses=True)
            worst, per = 0.0, {}
            for kind in ("lt", "lr"):
                for k, v in first_losses[kind].items():
                    e = rel_err(v, ref_first[kind][k])
                    per["%s.%s" % (kind, k)] = {"gpu": v, "oracle": ref_first[kind][k], "rel_err": e}
                    worst = max(worst, e)
            parity = {"what": "first-step losses (fresh weights) of the LT and LR steps at B=%d vs the CPU oracle (fp32)" % B,
                      "tolerance_rel": 1e-3, "max_rel_err": worst, "ok": bool(worst <= 1e-3), "losses": per}
            if n_cpu and extras:
                try:        # config #1 is quoted on the reference's CPU path: time the oracle's flow step beside the GPU number
                    extras[0]["cpu_baseline"] = cpu_flow_steps(256, 20, 2)
                except Exception as exc:      # a reported baseline must never take the bench line down
                    extras[0]["cpu_baseline"] = {"error": repr(exc)[:200]}
            if n_cpu:
                cpu_base = {"value": B / cpu_sec, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "oracle port (PyTorch CPU fp32), LT+LR step on %d-pose batches (the bench batch), %d "
                                      "timed steps after 1 warm-up" % (B, n_cpu)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: leg/torso + left/right lifter self-supervised training step, "
                                   "B=%d poses per GPU (N=%d rows), 17 joints" % (B, 2 * B),
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "cuda_graph": not args.no_graph,
                       "step": "LT and LR steps merged: one 4-network engine (every layer of both in the same GEMM launches), "
                               "%s GEMM launches per step, sampling flow evaluated once and %s" % (
                                   gemm_launches_per_step, "prefetched one step ahead" if not args.no_prefetch else "inline"),
                       "gemm_mode": "chain (one persistent kernel per pass, tile-level dependencies)" if MLP.USE_CHAIN
                                    else "layer-by-layer grouped launches",
                       "l2_policy": "no explicit flush: each step streams ~%.1f GB of activations/weights/gradients, "
                                    "far above the 126 MB L2" % (2 * 2048 * 1024 * 2 * 2 * 60 * (B / 1024) / 1e9),
                       "operands": "bf16 operands + bf16-stored activations, fp32 accumulate / master weights / losses",
                       "elevation_stats": "global" if (world == 1 or args.global_elevation_stats) else "local shard",
                       "grad_allreduce": ("none" if world == 1 else
                                          "push: wgrad epilogues store bf16 tiles into the owner rank's staging buffer over "
                                          "NVLink, sharded Adam, bf16 shadows stored into every rank" if args.grad_comm == "push"
                                          else "%s buckets over NCCL, overlapped with backward" % args.grad_comm),
                       "final_losses": final_losses},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": dict({"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                              "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
                              "traffic_note": "traffic = mean dram read+write bytes per chain launch from this round's ncu "
                                              "--set full capture of the same four launches at B=1024 "
                                              "(profiles/r02_gemm_traffic.json: 551 / 192 / 460 / 2219 MB); algorithmic operand "
                                              "bytes: %.1f MB per (2048 x 1024 x 1024) layer problem (A, W, out in bf16), i.e. "
                                              "~0.65 / 0.39 / 0.39 GB for the forward / pass-2 dgrad chains (measured traffic is "
                                              "BELOW it: layer outputs are consumed from L2) and ~1.45 GB for pass-1 dgrad + "
                                              "weight gradients (measured 1.5x: G / X slices re-read by the 4 tiles that share "
                                              "them)" % ((2048 * 1024 * 2 * 2 + 1024 * 1024 * 2) / 1e6),
                              "flops_per_launch": gemm_flops / n_gemm_launches,
                              "us_per_launch": ms_gemm * 1e3 / n_gemm_launches,
                              "launches_per_step": n_gemm_launches,
                              "kernel": "links::gemm_kernel<%s> (tcgen05 cta_group::2 / TMEM / TMA): the %d GEMM launches of one "
                                        "step as pure GEMM work on all SMs, timed in isolation" % (
                                            "true" if MLP.USE_CHAIN else "false", n_gemm_launches),
                              "ms_per_step_gemm_only": ms_gemm, "flops_per_step": gemm_flops, "in_step": in_step,
                              "frac_of_sustained_peak": achieved / peaks["bf16_sustained"],
                              "peak_source": "%s cuBLAS bf16 burst (kernel timed alone)" % peaks["source"]}, **sim),
            "parity": parity,
            "cpu_baseline": cpu_base,
            "strong_scaling": strong,
            "configs": extras,
        }
        emit(json.dumps(line))
    if world > 1:
        # captured graphs hold NCCL work; drop them before the communicator goes away.  Tearing the process group down
        # with captured collectives alive was observed to hang on exit, so leave without running destructors.
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
    ap.add_argument("--impl", default="links_b200", choices=["links_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="evaluate the sampling flow inside the step instead of one step ahead")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--dp-buckets", type=int, default=2, help="gradient buckets per network set under data parallelism")
    ap.add_argument("--nccl-ctas", type=int, default=0, help="CTAs (SMs) the NCCL all-reduce kernels may use (0: NCCL default)")
    ap.add_argument("--gemm-ctas", type=int, default=-1, help="GEMM grid cap under data parallelism (<= 0: no cap)")
    ap.add_argument("--grad-comm", default="push", choices=["push", "bf16", "fp32"],
                    help="data-parallel gradient exchange: push = reduce-scatter by peer stores fused into the weight-gradient "
                         "GEMM epilogues + sharded Adam + shadow all-gather by peer stores (NVLink, symmetric memory); "
                         "bf16 / fp32 = bucketed NCCL all-reduce of compressed / fp32 gradients")
    ap.add_argument("--global-elevation-stats", action="store_true",
                    help="elevation statistic (props.mean()/std()) over the global batch instead of each rank's shard")
    ap.add_argument("--eval-poses", type=int, default=1_250_000, help="poses per GPU of the config #5 eval run")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="only the headline config (no configs[] / strong_scaling)")
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
