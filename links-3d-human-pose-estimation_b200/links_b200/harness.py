"""Plain training / evaluation harness behind the drop-in scripts (replaces the pytorch_lightning 1.x harness of the
reference, which is not installable here -- SURVEY F4): same step order, same optimiser / scheduler settings, same
checkpoint file names, synthetic data when no dataset pickle is available.

Reference: train_leg_torso_lifter.py:61-121,376-398, train_left_right_lifter.py:59-119,541-560,
train_occlusion_models.py:81-142,547-570, eval_h36m.py:27-99."""
import os
import sys
import time

import torch

from . import init as INIT
from .occlusion import OCC_IN, OCC_NAMES, OCC_OUT, EvalRunner, OcclusionStep
from .data import ArrayLoader, DevicePrefetcher, loader_from_dataset
from .shard import shard_bounds
from .steps import LifterStep
from .synth import synth_poses

LR0, GAMMA = 2e-4, 0.95          # config.learning_rate, ExponentialLR(gamma=0.95) stepped per epoch


def add_common_args(p, batch=256, epochs=100):
    """Flags the reference does not have (the reference's own flags are declared by each script, verbatim)."""
    g = p.add_argument_group("links_b200 additions")
    g.add_argument("--synthetic", type=int, default=65536, help="number of synthetic training poses (no dataset is shipped)")
    g.add_argument("--batch", type=int, default=batch, help="global batch size (reference: config.BATCH_SIZE)")
    g.add_argument("--epochs", type=int, default=epochs, help="reference: config.N_epochs")
    g.add_argument("--steps", type=int, default=0, help="stop after this many optimiser steps (0 = run all epochs)")
    g.add_argument("--seed", type=int, default=0)
    g.add_argument("--weights-dir", default="models", help="where pretrained flows / lifters are read from and results saved")
    g.add_argument("--log-every", type=int, default=50)
    g.add_argument("--no-save", action="store_true")
    g.add_argument("--datafile", default=None, help="dataset pickle in the reference's format ({subject: {poses_2d, poses_3d}})")
    g.add_argument("--dataset", default="h36m", choices=["h36m", "mpi"], help="which dataset class reads --datafile")
    g.add_argument("--val", type=int, default=0, help="synthetic validation poses scored at every epoch end (0 = off)")
    g.add_argument("--random-init", action="store_true",
                   help="use seeded random weights for pretrained networks whose checkpoint is missing (benchmarks / "
                        "smoke runs); without it a missing checkpoint raises FileNotFoundError like the reference's torch.load")
    g.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    g.add_argument("--grad-comm", default="push", choices=["push", "bf16", "fp32"],
                   help="data-parallel gradient exchange of the lifter trainers: push = reduce-scatter by peer stores out of the "
                        "weight-gradient GEMM epilogues + sharded Adam (NVLink, symmetric memory); bf16 / fp32 = NCCL all-reduce")
    g.add_argument("--global-elevation-stats", action="store_true",
                   help="data parallel: props.mean() / props.std() over the global batch instead of each rank's shard")
    g.add_argument("--no-prefetch", action="store_true",
                   help="lifter trainers: draw the sampled poses inside the step instead of one step ahead")
    return p


def dist_setup():
    """One process per GPU under torchrun (NCCL over NVLink); single process otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = torch.distributed.group.WORLD
    return rank, world, pg


# Checkpoint names.  The reference's readers and writers do not agree with each other (e.g. train_full_pose_norm_flow.py:115
# writes models/norm_flow_sampling.pt, train_leg_torso_lifter.py:367 reads models/norm_flow_full_pose_with_sampling.pt), so
# every reader here accepts, in this order: the name the reference script READS, the name the producing reference script
# WRITES (so the shipped pipeline chains end to end), and the round-1 name of this repo.
CKPT = {
    "full_flow": ("norm_flow_full_pose_with_sampling.pt", "norm_flow_sampling.pt", "mpi_norm_flow_sampling.pt",
                  "full_pose_norm_flow.pt"),
    "full_flow_parts": ("mpi_norm_flow_sampling.pt", "norm_flow_sampling.pt", "norm_flow_full_pose_with_sampling.pt"),
    "leg_flow": ("best_lifting_models/no_sched_norm_flow_leg_weights_h36m_v5.pt", "mpi_norm_flow_legs_2.pt", "leg_norm_flow.pt"),
    "torso_flow": ("best_lifting_models/no_sched_norm_flow_torso_weights_h36m_v5.pt", "mpi_norm_flow_torso_2.pt",
                   "torso_norm_flow.pt"),
    "left_flow": ("no_sched_norm_flow_left_side_weights_h36m_v3_sampling.pt", "mpi_norm_flow_left_2.pt", "left_norm_flow.pt"),
    "right_flow": ("no_sched_norm_flow_right_side_weights_h36m_v3_sampling.pt", "mpi_norm_flow_right_2.pt", "right_norm_flow.pt"),
    "leg_lifter": ("legs_lifter.pt", "leg_lifter.pt"),                       # train_occlusion_models.py:530 / train_leg_torso_lifter.py:397
    "torso_lifter": ("torso_lifter.pt",),
    "left_lifter": ("left_lifter.pt", "final_best_left_lifter.pt", "left_side_lifter_final.pt"),    # eval_h36m.py:33, occlusion :532, LR trainer :560
    "right_lifter": ("right_lifter.pt", "final_best_right_lifter.pt", "right_side_lifter_final.pt"),
}


def load_state(path, fallback=None, allow_random=False):
    """Reference checkpoints are plain state dicts (torch.save(module.state_dict())).  `path`: one path or a list of
    candidates (first existing wins).  A missing checkpoint raises FileNotFoundError -- what the reference's torch.load
    does -- unless allow_random (--random-init) asks for seeded random weights."""
    paths = [path] if isinstance(path, str) else list(path)
    for q in paths:
        if os.path.exists(q):
            return {k: v.float() for k, v in torch.load(q, map_location="cpu").items()}
    if allow_random and fallback is not None:
        print("[links_b200] none of %s found: --random-init -> seeded random initialisation" % paths, file=sys.stderr)
        return fallback()
    raise FileNotFoundError("pretrained checkpoint not found (tried %s); train it with the producing script first, or "
                            "pass --random-init for seeded random weights" % ", ".join(paths))


def ckpt_paths(weights_dir, key):
    return [os.path.join(weights_dir, n) for n in CKPT[key]]


def save_module_state(module_cls, kwargs, params, path):
    """Save with the reference's full key set (unused LayerNorm / res_common entries included)."""
    m = module_cls(**kwargs)
    m.load_state_dict(params, strict=False)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(m.state_dict(), path)


class SyntheticLoader(ArrayLoader):
    """Epochs of shuffled batches of this rank's shard of a synthetic pose set, staged through pinned host memory."""

    def __init__(self, n, batch_global, rank, world, seed):
        x2d, gt = synth_poses(n, seed=1234 + seed)
        super().__init__(x2d, gt, batch_global, rank, world, seed)


def make_loader(args, rank, world):
    """The reference's dataset pickle when --datafile is given (H36M_Data / MPI_INF_3DHP_Dataset with normalize_head and
    ground-truth 2D, as train_leg_torso_lifter.py:376-386 builds it), otherwise synthetic poses."""
    if getattr(args, "datafile", None):
        from utils.helpers import normalize_head
        if args.dataset == "mpi":
            from utils.mpi_inf_3dhp_dataset_class import MPI_INF_3DHP_Dataset as DS
        else:
            from utils.h36m_dataset_class import H36M_Data as DS
        ds = DS(args.datafile, train=True, normalize_func=normalize_head, get_2dgt=True)
        return loader_from_dataset(ds, args.batch, rank, world, args.seed)
    return SyntheticLoader(args.synthetic, args.batch, rank, world, args.seed)


class Validator:
    """Epoch-end validation on the device (reference validation_step, train_leg_torso_lifter.py:286-337 /
    train_left_right_lifter.py:437-520): lift the validation poses with the CURRENT lifters, then PA-MPJPE ('best'),
    scale-matched MPJPE, PCK@150 and AUC.  The reference scores PA-MPJPE with a per-pose numpy loop on the host; here
    the whole validation set is one fused lift+score pass per chunk plus two threshold-count launches."""

    def __init__(self, kind, step, n_val, seed, depth, rank, world, pg, chunk=16384):
        x2d, gt = synth_poses(n_val, seed=4321 + seed)
        b, e = shard_bounds(n_val, rank, world, multiple=1)
        dev = step.device
        self.x, self.gt = torch.from_numpy(x2d[b:e]).to(dev), torch.from_numpy(gt[b:e]).to(dev)
        self.step, self.pg, self.world = step, pg, world
        self.chunk = min(chunk, max(e - b, 1))
        step.mlp.zero_sync_master()            # push mode: each rank only keeps its own rows of the big layers current
        self.ev = EvalRunner(kind, [step.mlp.state_dict(s) for s in range(2)], chunk=self.chunk, depth=depth,
                             choice="right", process_group=pg)

    def run(self):
        from utils.metrics_batch import Metrics as mb
        self.step.mlp.zero_sync_master()
        self.ev.load_lifters([self.step.mlp.state_dict(s) for s in range(2)])
        self.ev.reset()
        preds = [self.ev.run_chunk(self.x[i:i + self.chunk], self.gt[i:i + self.chunk], want_pred=True)
                 for i in range(0, self.x.shape[0], self.chunk)]
        pred = torch.cat(preds, dim=0)
        out = self.ev.result()
        # PCK / AUC are means over poses of per-pose hit ratios: reduce (ratio * count, count) pairs over the ranks
        # (validation shards need not be equal)
        n_loc = float(self.gt.shape[0])
        extra = torch.stack((mb().PCK(self.gt, pred, num_joints=17, root_joint=0).double() * n_loc,
                             mb().AUC(self.gt, pred, num_joints=17, root_joint=0).double() * n_loc,
                             torch.tensor(n_loc, dtype=torch.float64, device=pred.device)))
        if self.world > 1:
            torch.distributed.all_reduce(extra, group=self.pg)
        extra = extra[:2] / extra[2]
        return {"pa": out["pa_mpjpe"], "mpjpe_scaled": out["n_mpjpe"], "pck": extra[0].item(), "auc": extra[1].item()}


def run_training(step, loader, args, rank, feed, validator=None):
    """Epoch loop: feed(step, batch, ...) uploads inputs + draws, then the step runs on the device.

    The whole step is captured into ONE CUDA graph after the first (eager) step and replayed from then on (--no-graph
    keeps eager launches): the static input buffers are refilled by `feed` on the same stream before every replay, the
    learning rate lives in a device word (step.set_lr), so ExponentialLR needs no re-capture.
    Steps with sampling prefetch (LifterStep(cfg prefetch_sample)) draw the poses of batch i+1 while batch i trains:
    feed(..., sampling=True) loads the NEXT batch's sampling inputs, feed(..., draws=True) this step's random draws."""
    n_steps, t0 = 0, time.time()
    lr = LR0
    use_graph = not getattr(args, "no_graph", False)
    prefetch = bool(getattr(step, "prefetch", False))
    graph = None

    def run_step():
        nonlocal graph
        if graph is not None:
            graph.replay()
        elif use_graph and n_steps >= 1 and hasattr(step, "capture"):
            graph = step.capture(warmup=0)          # plans exist after the first eager step
            graph.replay()
        else:
            step.step()

    for epoch in range(args.epochs):
        step.set_lr(lr)
        it = iter(DevicePrefetcher(loader, step.device))     # batch i+1 crosses PCIe while step i runs
        pending = None
        if prefetch:
            first = next(it, None)
            if first is None:
                break
            feed(step, first, sampling=True, draws=False)
            step.prime()                                     # poses of the first batch
            pending = first
        while True:
            xb = next(it, None)
            if prefetch:
                if pending is None:
                    break
                if xb is not None:
                    feed(step, xb, sampling=True, draws=False)     # sampled while this step trains on `pending`
                feed(step, pending, sampling=False, draws=True)
                pending = xb
            else:
                if xb is None:
                    break
                feed(step, xb, sampling=True, draws=True)
            run_step()
            n_steps += 1
            if rank == 0 and n_steps % args.log_every == 0:
                torch.cuda.synchronize()
                d = step.loss_dict()
                flat = d if not isinstance(next(iter(d.values())), dict) else \
                    {"%s.%s" % (k, kk): vv for k, v in d.items() for kk, vv in v.items()}
                print("epoch %d step %d  %s  (%.0f poses/s)" % (epoch, n_steps, " ".join("%s=%.5f" % kv for kv in flat.items()),
                                                                n_steps * args.batch / (time.time() - t0)), flush=True)
            if args.steps and n_steps >= args.steps:
                break
        if validator is not None:
            v = validator.run()
            if rank == 0:
                print("epoch %d validation  %s" % (epoch, " ".join("%s=%.4f" % kv for kv in v.items())), flush=True)
        if args.steps and n_steps >= args.steps:
            break
        lr *= GAMMA                                   # training_epoch_end: ExponentialLR.step()
    torch.cuda.synchronize()
    return n_steps


def lifter_feed(gen_dev):
    def feed(step, xb, sampling=True, draws=True):
        if sampling:
            assert xb.shape[0] == step.B
            step.x.copy_(xb, non_blocking=True)
            step.noise.normal_(generator=gen_dev)        # add_noise eps (utils/helpers.py:298-308)
        if draws:
            step.eps_x.normal_(generator=gen_dev)        # elevation draw (train_leg_torso_lifter.py:169-171)
            step.u_y.uniform_(generator=gen_dev)         # azimuth draw (:176)
    return feed


def train_lifters(kind, args):
    rank, world, pg = dist_setup()
    cfg = dict(depth=args.translation, weight_bl=args.bl, weight_2d=args.rep2d, weight_3d=args.rot3d,
               weight_likeli=args.likelihood, weight_velocity=args.velocity)
    wd = args.weights_dir
    if kind == "lt":
        nj, names = (7, 10), ("leg_lifter.pt", "torso_lifter.pt")
        flow_keys = ("leg_flow", "torso_flow")
    else:
        nj, names = (11, 11), ("left_side_lifter_final.pt", "right_side_lifter_final.pt")
        flow_keys = ("left_flow", "right_flow")
    nets = [INIT.init_lifter_params(n, 11 + i + args.seed) for i, n in enumerate(nj)]
    rnd = getattr(args, "random_init", False)
    flows = [load_state(ckpt_paths(wd, f), lambda C=2 * n, s=41 + i: INIT.init_flow_params(C, s), rnd)
             for i, (f, n) in enumerate(zip(flow_keys, nj))]
    full = load_state(ckpt_paths(wd, "full_flow"), lambda: INIT.init_flow_params(34, 40), rnd)
    loader = make_loader(args, rank, world)
    cfg["prefetch_sample"] = not getattr(args, "no_prefetch", False)
    cfg["grad_comm"] = getattr(args, "grad_comm", "push")
    cfg["global_elevation_stats"] = getattr(args, "global_elevation_stats", False)
    try:
        step = LifterStep(kind, loader.batch, nets, flows, full, cfg=cfg, process_group=pg)
    except Exception as e:  # noqa: BLE001
        if not (world > 1 and cfg["grad_comm"] == "push"):
            raise
        print("[links_b200] push mode unavailable (%s); using bf16 NCCL gradient buckets" % e, file=sys.stderr)
        cfg["grad_comm"] = "bf16"
        step = LifterStep(kind, loader.batch, nets, flows, full, cfg=cfg, process_group=pg)
    gen_dev = torch.Generator(device="cuda").manual_seed(args.seed * 7919 + rank)
    validator = Validator(kind, step, args.val, args.seed, args.translation, rank, world, pg) if args.val else None
    n = run_training(step, loader, args, rank, lifter_feed(gen_dev), validator)
    step.mlp.zero_sync_master()                # collective: complete fp32 masters on every rank before rank 0 saves them
    if rank == 0 and not args.no_save:
        from utils import models_def as MD
        cls = (MD.Leg_Lifter, MD.Torso_Lifter) if kind == "lt" else (MD.Left_Right_Lifter, MD.Left_Right_Lifter)
        for s in range(2):
            save_module_state(cls[s], dict(use_batchnorm=False, num_joints=nj[s], use_dropout=False, d_rate=0.25),
                              step.mlp.state_dict(s), os.path.join(wd, names[s]))
    return n, step


OCC_FILES = {"left_arm": "left_arm_estimator.pt", "right_arm": "right_arm_estimator.pt", "left_leg": "left_leg_estimator.pt",
             "right_leg": "right_leg_estimator.pt", "left": "left_side_estimator.pt", "right": "right_side_estimator.pt",
             "both_legs": "both_legs_estimator.pt", "torso": "torso_estimator.pt"}


def train_occlusion(args):
    rank, world, pg = dist_setup()
    wd = args.weights_dir
    rnd = getattr(args, "random_init", False)
    lifters = [load_state(ckpt_paths(wd, "leg_lifter"), lambda: INIT.init_lifter_params(7, 11), rnd),
               load_state(ckpt_paths(wd, "torso_lifter"), lambda: INIT.init_lifter_params(10, 12), rnd)]
    preds = {n: INIT.init_predictor_params(OCC_IN[n] // 3, OCC_OUT[n], 100 + i + args.seed) for i, n in enumerate(OCC_NAMES)}
    loader = make_loader(args, rank, world)
    step = OcclusionStep(loader.batch, lifters, preds, cfg=dict(depth=args.translation), process_group=pg)
    gen_dev = torch.Generator(device="cuda").manual_seed(args.seed * 7919 + rank)

    def feed(st, xb, sampling=True, draws=True):
        st.x.copy_(xb, non_blocking=True)
        st.u_y[0].uniform_(generator=gen_dev)        # Ry augmentation draws (train_occlusion_models.py:213-217,256-260)
        st.u_y[1].uniform_(generator=gen_dev)
    validator = None
    if args.val:                                # validation_step of the occlusion script (:316-509), on the device
        from .occ_assembly import OcclusionValidator
        lr = [load_state(ckpt_paths(wd, "left_lifter")[1:], lambda: INIT.init_lifter_params(11, 13), rnd),
              load_state(ckpt_paths(wd, "right_lifter")[1:], lambda: INIT.init_lifter_params(11, 14), rnd)]
        ov = OcclusionValidator.from_params({"legs": lifters[0], "torso": lifters[1], "left": lr[0], "right": lr[1]},
                                            {n: step.mlp.state_dict(s) for s, n in enumerate(OCC_NAMES)},
                                            depth=args.translation, device=step.device)
        x2d, gt = synth_poses(args.val, seed=4321 + args.seed)
        b, e = shard_bounds(args.val, rank, world, multiple=1)
        vx, vg = torch.from_numpy(x2d[b:e]).to(step.device), torch.from_numpy(gt[b:e]).to(step.device)

        class _V:
            def run(self_inner):
                ov.load_predictors({n: step.mlp.state_dict(s) for s, n in enumerate(OCC_NAMES)})
                out = ov.run(vx, vg)
                if world > 1:                   # per-pose means: reduce (mean * count, count) over the ranks
                    n_loc = float(vx.shape[0])
                    t = torch.tensor([v * n_loc for v in out.values()] + [n_loc], dtype=torch.float64, device=step.device)
                    torch.distributed.all_reduce(t, group=pg)
                    out = dict(zip(out, (t[:-1] / t[-1]).tolist()))
                return out
        validator = _V()
    n = run_training(step, loader, args, rank, feed, validator)
    if rank == 0 and not args.no_save:
        from utils import models_def as MD
        classes = {"left_arm": MD.Occluded_Limb_Predictor, "right_arm": MD.Occluded_Limb_Predictor,
                   "left_leg": MD.Occluded_Limb_Predictor, "right_leg": MD.Occluded_Limb_Predictor,
                   "left": MD.Occluded_Left_Right_Predictor, "right": MD.Occluded_Left_Right_Predictor,
                   "both_legs": MD.Occluded_Legs_Predictor, "torso": MD.Occluded_Torso_Predictor}
        for s, name in enumerate(OCC_NAMES):
            key = name if name in OCC_FILES else name.replace("_side", "")
            save_module_state(classes[key], dict(use_batchnorm=False, num_joints=OCC_IN[name] // 3), step.mlp.state_dict(s),
                              os.path.join(wd, "occlusion_model_weights", OCC_FILES[key]))
    return n, step


def evaluate(args):
    """eval_h36m.py:27-99 on a (sharded) synthetic test set: prints PA-MPJPE ('best') and N-MPJPE."""
    rank, world, pg = dist_setup()
    wd = args.weights_dir
    rnd = getattr(args, "random_init", False)
    lifters = [load_state(ckpt_paths(wd, "left_lifter"), lambda: INIT.init_lifter_params(11, 13), rnd),
               load_state(ckpt_paths(wd, "right_lifter"), lambda: INIT.init_lifter_params(11, 14), rnd)]
    if getattr(args, "datafile", None):          # eval_h36m.py:41: test subjects, fixed-scale normalisation
        from utils.h36m_dataset_class import H36M_Data
        from utils.helpers import normalize_head_test
        ds = H36M_Data(args.datafile, train=False, normalize_func=normalize_head_test, get_2dgt=True, subjects=['S9', 'S11'])
        x2d, gt = ds.data["poses_2d"].astype("float32"), ds.data["poses_3d"].astype("float32")
    else:
        x2d, gt = synth_poses(args.synthetic, seed=4321 + args.seed)
    b, e = shard_bounds(x2d.shape[0], rank, world, multiple=1)
    ev = EvalRunner("lr", lifters, chunk=args.chunk, choice="right", process_group=pg)
    xs, gs = torch.from_numpy(x2d[b:e]).pin_memory(), torch.from_numpy(gt[b:e]).pin_memory()
    for i in range(0, e - b, args.chunk):
        ev.run_chunk(xs[i:i + args.chunk].cuda(non_blocking=True), gs[i:i + args.chunk].cuda(non_blocking=True))
    out = ev.result()
    if rank == 0:
        print("PA-MPJPE: %.4f" % out["pa_mpjpe"])
        print("N-MPJPE: %.4f" % out["n_mpjpe"])
    return out
