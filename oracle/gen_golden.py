"""Generate tests/golden/*.npz by running the REFERENCE's own utils (build container only).

Test infrastructure.  Run from the repo root:  ``python -m oracle.gen_golden``.

Imports ``/root/reference/utils/{models_def,helpers,rotation_conversions,metrics_batch,
metrics}.py`` (with a stub ``pytorch_lightning``: models_def.py:2 imports it and never uses
it), asserts the oracle restatement reproduces them on seeded inputs, and freezes the
reference's outputs as small fixtures.  ``/root/reference`` does not exist on the GPU box,
so only the committed ``.npz`` files travel.  Network weights are not stored (59 MB per
lifter): fixtures hold the seed, and ``oracle.nets.init_*`` regenerates them bit-exactly
(torch CPU mt19937 generator).

FrEIA is not importable, so flow fixtures are produced by the (unpinned) restatement and
are regression vectors plus known-answer checks, not reference outputs.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "links-3d-human-pose-estimation_b200"))


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not present (only exists in the build container)" % REF)
    sys.modules.setdefault("pytorch_lightning", types.ModuleType("pytorch_lightning"))
    import importlib.util
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # metrics.py:134 `is not 'best'` SyntaxWarning
        for name in ("models_def", "helpers", "rotation_conversions", "metrics_batch", "metrics"):
            spec = importlib.util.spec_from_file_location("linksref_" + name, os.path.join(REF, "utils", name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods[name] = m
    return mods


def t(a):
    return a.detach().cpu().numpy()


def helpers_extra(ref):
    """Outputs of the reference's remaining utils/helpers.py functions (the ones no step body calls but a user of the
    module may: alternative splits, occluded-pose assembly, part bone lengths, fixed-scale normalisers, part
    projections, latent interpolation, occlusion masks) on seeded inputs -> tests/golden/helpers_extra.npz."""
    import random
    H = ref["helpers"]
    g = torch.Generator().manual_seed(123)
    out = {}
    a34 = torch.randn(6, 34, generator=g)
    out["a34"] = t(a34)
    l, r = H.split_data_left_right_v2(a34)
    out["split_v2_left"], out["split_v2_right"] = t(l), t(r)
    ln, rn = H.split_data_left_right_numpy(a34.numpy())
    out["split_np_left"], out["split_np_right"] = ln, rn
    a68 = torch.randn(4, 68, generator=g)
    out["a68"] = t(a68)
    l, r = H.temporal_split_data_left_right(a68)
    out["split_temporal_left"], out["split_temporal_right"] = t(l), t(r)
    occ, vis = torch.randn(5, 18, generator=g), torch.randn(5, 33, generator=g)
    out["occ_part"], out["vis_part"] = t(occ), t(vis)
    for side in ("right", "left"):
        out["combine_occluded_" + side] = t(H.combine_left_right_occluded_3d(occ, vis, side))
    legs, lr = torch.randn(5, 21, generator=g), torch.randn(5, 33, generator=g)
    out["legs3d"], out["lr3d"] = t(legs), t(lr)
    out["bones_legs"] = t(H.get_bone_lengths_legs(legs))
    out["bones_left_right"] = t(H.get_bone_lengths_left_right(lr))
    raw2d = np.random.RandomState(9).normal(size=(7, 34)) * 120 + 400
    out["raw2d"] = raw2d
    for fn in ("normalize_head_test_mpi_chest", "normalize_head_test_mpi_vnect", "normalize_head_test_temporal"):
        out[fn] = getattr(H, fn)(raw2d.copy())
    lat = torch.randn(8, 34, generator=g)
    out["latent"] = t(lat)
    out["interp_0.3"] = t(H.interpolate_gaussian_batch(lat, 0.3))
    for fn, w in (("perspective_projection_legs", 21), ("perspective_projection_torso", 30),
                  ("perspective_projection_left_right", 33)):
        p = torch.randn(5, w, generator=g)
        p[:, 2 * w // 3:] = p[:, 2 * w // 3:].abs() + 4.0
        out[fn + "_in"], out[fn + "_out"] = t(p), t(getattr(H, fn)(p))
    random.seed(77)
    out["occlusion_create"] = t(H.occlusion_create(a34))
    np.savez_compressed(os.path.join(OUT, "helpers_extra.npz"), **out)
    print("helpers_extra.npz written (%d arrays)" % len(out))


MODULE_CASES = [("PoseDiscriminator", dict(num_joints=16)), ("DepthAngleEstimator", dict(num_joints=16)),
                ("Leg_Lifter", dict(num_joints=7, d_rate=0.25)), ("Torso_Lifter", dict(num_joints=10, d_rate=0.25)),
                ("Left_Right_Lifter", dict(num_joints=11, d_rate=0.25)), ("Occluded_Limb_Predictor", dict(num_joints=14)),
                ("Occluded_Legs_Predictor", dict(num_joints=10)), ("Occluded_Torso_Predictor", dict(num_joints=7)),
                ("Occluded_Left_Right_Predictor", dict(num_joints=11)), ("res_block", dict())]


def module_contract(ref):
    """state_dict key names + shapes, constructor signatures and default arguments of every class in the reference's
    utils/models_def.py -> tests/golden/module_contract.json (the drop-in boundary)."""
    import inspect
    import json
    M = ref["models_def"]
    out = {}
    for name, kw in MODULE_CASES:
        cls = getattr(M, name)
        sig = inspect.signature(cls.__init__)
        mod = cls(use_batchnorm=False, **kw)
        out[name] = {"kwargs": kw,
                     "signature": [[p.name, None if p.default is inspect._empty else p.default]
                                   for p in list(sig.parameters.values())[1:]],
                     "state_dict": [[k, list(v.shape)] for k, v in mod.state_dict().items()]}
    with open(os.path.join(OUT, "module_contract.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("module_contract.json written (%d classes)" % len(out))


def cli_contract():
    """Command-line flags of the reference's scripts (option strings, type, default), read from their argparse
    sections with the ast module (the scripts cannot be imported: module-level parse_args / wandb / .cuda()) ->
    tests/golden/cli_contract.json."""
    import ast
    import json
    out = {}
    for script in ("train_leg_torso_lifter.py", "train_left_right_lifter.py", "train_occlusion_models.py",
                   "train_full_pose_norm_flow.py", "train_leg_torso_left_right_norm_flow.py", "eval_h36m.py"):
        tree = ast.parse(open(os.path.join(REF, script)).read())
        flags = []
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
                opts = [a.value for a in node.args if isinstance(a, ast.Constant)]
                kw = {k.arg: (k.value.id if isinstance(k.value, ast.Name) else ast.literal_eval(k.value)) for k in node.keywords
                      if k.arg in ("type", "default")}
                flags.append({"options": opts, "type": kw.get("type"), "default": kw.get("default")})
        out[script] = flags
    with open(os.path.join(OUT, "cli_contract.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("cli_contract.json written:", {k: len(v) for k, v in out.items()})


def script_functions():
    """Module-level helper functions that live inside the (non-importable) scripts: the FunctionDef is cut out of the
    script's AST and executed on its own, so the golden outputs still come from the reference's code ->
    tests/golden/script_funcs.npz.  Currently: combine_pose_and_limb (train_occlusion_models.py:67-78)."""
    import ast
    tree = ast.parse(open(os.path.join(REF, "train_occlusion_models.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "combine_pose_and_limb"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_occlusion_models.py", "exec"), ns)
    pose = torch.arange(5 * 42, dtype=torch.float32).reshape(5, 42)
    limb = 1000 + torch.arange(5 * 9, dtype=torch.float32).reshape(5, 9)
    out = {"pose": t(pose), "limb": t(limb)}
    for which in ("ll", "rl", "la", "ra"):
        out["combine_" + which] = t(ns["combine_pose_and_limb"](pose, limb, which))
    np.savez_compressed(os.path.join(OUT, "script_funcs.npz"), **out)
    print("script_funcs.npz written")


def metrics_np_extra(ref):
    """utils/metrics.py (the per-pose numpy class) beyond what metrics.npz holds: mpjpe flag combinations, PCK, and the
    full (d, Z, tform) output of procrustes for every scaling / reflection setting -> tests/golden/metrics_np_extra.npz."""
    from links_b200.synth import synth_poses, synth_pred_3d
    m = ref["metrics"].Metrics()
    _, gt = synth_poses(6, seed=71)
    pred = synth_pred_3d(gt, seed=72, noise_mm=35.0, scale=0.013, mirror_frac=0.5)
    gt, pred = gt.astype(np.float64), pred.astype(np.float64)
    out = {"gt": gt, "pred": pred}
    for sc in (False, True):
        for ma in (False, True):
            out["mpjpe_s%d_m%d" % (sc, ma)] = np.array([m.mpjpe(gt[i:i + 1], pred[i:i + 1], scale=sc, mean_align=ma)
                                                         for i in range(6)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for refl in (False, True, "best"):
            out["pck_%s" % refl] = np.array([m.PCK(gt[i:i + 1], pred[i:i + 1], reflection=refl) for i in range(6)])
            for scaling in (True, False):
                d, Z, tf = m.procrustes(gt[0].reshape(3, 17).T, pred[0].reshape(3, 17).T, scaling=scaling, reflection=refl)
                tag = "proc_%s_%d_" % (refl, scaling)
                out[tag + "d"], out[tag + "Z"] = np.float64(d), Z
                out[tag + "rot"], out[tag + "scale"], out[tag + "trans"] = tf["rotation"], np.float64(tf["scale"]), tf["translation"]
    np.savez_compressed(os.path.join(OUT, "metrics_np_extra.npz"), **out)
    print("metrics_np_extra.npz written (%d arrays)" % len(out))


def reference_training_steps(ref):
    """Run the reference's OWN training_step code: the method is cut out of each script's AST (the scripts themselves
    cannot be imported) and executed on a stand-in `self` that carries the reference's networks (utils/models_def.py,
    importable) and -- the one substitution -- the oracle flow in place of the absent FrEIA modules.  Losses and a few
    gradients go to tests/golden/ref_steps.npz; tests/test_oracle_steps_pinned.py replays the same seeded inputs through
    oracle/steps.py.  RNG: the step draws randn_like(z) [B,34], normal(0,1) [N,1], rand [N,1] in that order (occlusion:
    rand [B,1] twice); the test re-seeds and draws the same sequence."""
    import ast
    from types import SimpleNamespace
    from oracle import flow as OF, nets as ON, steps as OS
    from links_b200.synth import synth_poses
    H, M, RC = ref["helpers"], ref["models_def"], ref["rotation_conversions"]

    def method(script, cls_name, name):
        tree = ast.parse(open(os.path.join(REF, script)).read())
        cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
        fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
        ns = {k: getattr(H, k) for k in dir(H) if not k.startswith("_")}
        ns.update(torch=torch, np=np, euler_angles_to_matrix=RC.euler_angles_to_matrix, SimpleNamespace=SimpleNamespace)
        exec(compile(ast.Module(body=[fn], type_ignores=[]), script, "exec"), ns)
        return ns

    class Flow:
        def __init__(self, params):
            self.p = params

        def __call__(self, x, rev=False):
            return OF.inn_forward(x, self.p, rev=rev)

    class Opt:
        def zero_grad(self):
            pass

        def step(self):
            pass

    def module(cls, nj, params):
        m = getattr(M, cls)(use_batchnorm=False, num_joints=nj, use_dropout=False, d_rate=0.25)
        m.load_state_dict(params, strict=False)
        return m

    def base_self(n_opt):
        return SimpleNamespace(optimizers=lambda: [Opt() for _ in range(n_opt)], device=torch.device("cpu"),
                               manual_backward=lambda loss: loss.backward(), losses=SimpleNamespace(), log=lambda *a, **k: None,
                               losses_mean=SimpleNamespace())

    cfg = SimpleNamespace(use_elevation=True, depth=10.0, weight_bl=50.0, weight_2d=1.0, weight_3d=1.0, weight_likeli=1.0,
                          weight_velocity=1.0)
    out = {}
    B = 16
    x2d, _ = synth_poses(B, seed=91)
    x = torch.from_numpy(x2d)
    out["x"] = x2d
    full = OF.init_flow_params(34, 40, perturb=0.3)
    cuda_orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # the step bodies call .cuda() on CPU tensors
    try:
        # ---- leg / torso step (train_leg_torso_lifter.py:123-284)
        ns = method("train_leg_torso_lifter.py", "LitLifter", "training_step")
        ns["config"] = cfg
        leg, torso = ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)
        s = base_self(2)
        s.full_inn_2d, s.leg_inn_2d, s.torso_inn_2d = Flow(full), Flow(OF.init_flow_params(14, 41, perturb=0.3)), \
            Flow(OF.init_flow_params(20, 42, perturb=0.3))
        s.legs_lifter, s.torso_lifter = module("Leg_Lifter", 7, leg), module("Torso_Lifter", 10, torso)
        from oracle import geometry as OG
        s.bone_relations_mean = torch.tensor(OG.BONE_REL_MPI, dtype=torch.float32)
        torch.manual_seed(1001)
        ns["training_step"](s, {"p2d_gt": x.clone()}, 0)
        for k, v in s.losses.__dict__.items():
            out["lt_" + k] = np.float64(v.item())
        out["lt_dW_leg_upscale"] = t(s.legs_lifter.upscale.weight.grad)
        out["lt_dW_torso_angles"] = t(s.torso_lifter.angles.weight.grad)
        # ---- left / right step (train_left_right_lifter.py:121-435)
        ns = method("train_left_right_lifter.py", "LitLifter", "training_step")
        ns["config"] = cfg
        left, right = ON.init_lifter_params(11, 13), ON.init_lifter_params(11, 14)
        s = base_self(2)
        s.full_inn2d, s.left_inn2d, s.right_inn2d = Flow(full), Flow(OF.init_flow_params(22, 43, perturb=0.3)), \
            Flow(OF.init_flow_params(22, 44, perturb=0.3))
        s.left_lifter, s.right_lifter = module("Left_Right_Lifter", 11, left), module("Left_Right_Lifter", 11, right)
        s.bone_relations_mean = torch.tensor(OG.BONE_REL_H36M, dtype=torch.float32)
        torch.manual_seed(1002)
        ns["training_step"](s, {"p2d_gt": x.clone()}, 0)
        for k, v in s.losses.__dict__.items():
            out["lr_" + k] = np.float64(v.item())
        out["lr_dW_left_upscale"] = t(s.left_lifter.upscale.weight.grad)
        out["lr_dW_right_downscale"] = t(s.right_lifter.downscale.weight.grad)
        # ---- occlusion step (train_occlusion_models.py:144-314)
        ns = method("train_occlusion_models.py", "Limb_Predictor", "training_step")
        ns["config"] = cfg
        s = base_self(8)
        s.leg_lifter, s.torso_lifter = module("Leg_Lifter", 7, leg), module("Torso_Lifter", 10, torso)
        s.left_lifter, s.right_lifter = module("Left_Right_Lifter", 11, left), module("Left_Right_Lifter", 11, right)
        attr = {"left_arm": "left_arm_predictor", "right_arm": "right_arm_predictor", "left_leg": "left_leg_predictor",
                "right_leg": "right_leg_predictor", "left_side": "left_predictor", "right_side": "right_predictor",
                "both_legs": "both_legs_predictor", "torso": "torso_predictor"}
        cls = {"left_arm": "Occluded_Limb_Predictor", "right_arm": "Occluded_Limb_Predictor", "left_leg": "Occluded_Limb_Predictor",
               "right_leg": "Occluded_Limb_Predictor", "left_side": "Occluded_Left_Right_Predictor",
               "right_side": "Occluded_Left_Right_Predictor", "both_legs": "Occluded_Legs_Predictor", "torso": "Occluded_Torso_Predictor"}
        nin = {"left_arm": 14, "right_arm": 14, "left_leg": 14, "right_leg": 14, "left_side": 11, "right_side": 11,
               "both_legs": 11, "torso": 7}
        nout = {"left_arm": 9, "right_arm": 9, "left_leg": 9, "right_leg": 9, "left_side": 18, "right_side": 18,
                "both_legs": 18, "torso": 30}
        for i, n in enumerate(OS.OCC_NAMES):
            m = getattr(M, cls[n])(use_batchnorm=False, num_joints=nin[n])
            m.load_state_dict(ON.init_predictor_params(nin[n], nout[n], 100 + i), strict=False)
            setattr(s, attr[n], m)
        torch.manual_seed(1003)
        ns["training_step"](s, {"p2d_gt": x.clone()}, 0)
        for k, v in s.losses.__dict__.items():
            out["occ_" + k] = np.float64(v.item())
        out["occ_dW_torso_downscale"] = t(s.torso_predictor.downscale.weight.grad)
        out["occ_dW_left_upscale"] = t(s.left_predictor.upscale.weight.grad)
        # ---- validation steps (train_leg_torso_lifter.py:286-337, train_occlusion_models.py:317-509): per-pose numpy
        #      PA-MPJPE loops + metrics_batch, exactly as the scripts run them
        xv2d, gtv = synth_poses(12, seed=92)
        out["val_x"], out["val_gt"] = xv2d, gtv
        val_batch = {"p2d_gt": torch.from_numpy(xv2d), "poses_3d": torch.from_numpy(gtv)}
        wandb_stub = SimpleNamespace(log=lambda *a, **k: None)
        cfg.use_gt = True
        ns = method("train_leg_torso_lifter.py", "LitLifter", "validation_step")
        ns.update(config=cfg, wandb=wandb_stub, mb=ref["metrics_batch"].Metrics)
        s = base_self(2)
        s.legs_lifter, s.torso_lifter = module("Leg_Lifter", 7, leg), module("Torso_Lifter", 10, torso)
        s.metrics, s.current_epoch = ref["metrics"].Metrics(), 0
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns["validation_step"](s, val_batch, 0)
        for k in ("pa", "mpjpe_scaled", "auc", "pck"):
            out["ltval_" + k] = np.float64(getattr(s.losses, k))
        ns = method("train_left_right_lifter.py", "LitLifter", "validation_step")      # :437-511, both combine choices
        ns.update(config=cfg, wandb=wandb_stub, mb=ref["metrics_batch"].Metrics)
        s = base_self(2)
        s.left_lifter, s.right_lifter = module("Left_Right_Lifter", 11, left), module("Left_Right_Lifter", 11, right)
        s.metrics, s.current_epoch = ref["metrics"].Metrics(), 0
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns["validation_step"](s, val_batch, 0)
        for k in ("pa_left", "pa_right", "mpjpe_scaled_left", "mpjpe_scaled_right"):
            out["lrval_" + k] = np.float64(getattr(s.losses, k))
        # ---- eval_h36m.py:46-97: the script's own top-level statements from `metrics = Metrics()` to the prints
        tree_e = ast.parse(open(os.path.join(REF, "eval_h36m.py")).read())
        first = [n.lineno for n in tree_e.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "metrics"][0]
        body = [n for n in tree_e.body if n.lineno >= first]
        ns_e = {k: getattr(H, k) for k in dir(H) if not k.startswith("_")}
        ns_e.update(torch=torch, np=np, Metrics=ref["metrics"].Metrics, mb=ref["metrics_batch"].Metrics,
                    left_lifter=module("Left_Right_Lifter", 11, left), right_lifter=module("Left_Right_Lifter", 11, right),
                    poses_2d=torch.from_numpy(xv2d), poses_3d=torch.from_numpy(gtv), print=lambda *a, **k: None)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exec(compile(ast.Module(body=body, type_ignores=[]), "eval_h36m.py", "exec"), ns_e)
        out["evalh36m_pa"], out["evalh36m_mpjpe_scaled"] = np.float64(ns_e["pa"]), np.float64(ns_e["mpjpe_scaled"])
        # ---- flow trainers: the body of the scripts' inner batch loop, up to the last optimizer.step()
        def loop_body(script):
            tr = ast.parse(open(os.path.join(REF, script)).read())
            outer = [n for n in tr.body if isinstance(n, ast.For) and getattr(n.target, "id", "") == "epoch"][0]
            inner = [n for n in outer.body if isinstance(n, ast.For)][0]
            last = max(i for i, n in enumerate(inner.body) if isinstance(n, ast.Expr) and isinstance(n.value, ast.Call)
                       and getattr(n.value.func, "attr", "") == "step")
            return compile(ast.Module(body=inner.body[:last + 1], type_ignores=[]), script, "exec")

        def flow_ns(extra):
            d = {k: getattr(H, k) for k in dir(H) if not k.startswith("_")}
            d.update(torch=torch, np=np, losses=SimpleNamespace(), sample={"p2d_gt": x.clone()})
            d.update(extra)
            return d
        fp = OS.params_require_grad(OF.init_flow_params(34, 45, perturb=0.3))
        ns_f = flow_ns(dict(inn_2d=Flow(fp), optimizer=Opt()))
        torch.manual_seed(1004)
        exec(loop_body("train_full_pose_norm_flow.py"), ns_f)
        for k, v in ns_f["losses"].__dict__.items():
            out["flowtrain_" + k] = np.float64(v.item())
        out["flowtrain_dW"] = t(fp["module_list.2.subnet.2.weight"].grad)
        parts = {n: OS.params_require_grad(OF.init_flow_params(w, 60 + i, perturb=0.3))
                 for i, (n, w) in enumerate((("legs", 14), ("torso", 20), ("left", 22), ("right", 22)))}
        ns_p = flow_ns(dict(inn_2d_legs_split=Flow(parts["legs"]), inn_2d_torso_split=Flow(parts["torso"]),
                            inn_2d_left_split=Flow(parts["left"]), inn_2d_right_split=Flow(parts["right"]),
                            full_pose_inn2d=Flow(full), left_split_opt=Opt(), right_split_opt=Opt(), leg_optimizer=Opt(),
                            torso_optimizer=Opt()))
        torch.manual_seed(1005)
        exec(loop_body("train_leg_torso_left_right_norm_flow.py"), ns_p)
        for k, v in ns_p["losses"].__dict__.items():
            out["partflow_" + k] = np.float64(v.item())
        out["partflow_dW_left"] = t(parts["left"]["module_list.1.subnet.0.weight"].grad)
        ns = method("train_occlusion_models.py", "Limb_Predictor", "validation_step")
        tree = ast.parse(open(os.path.join(REF, "train_occlusion_models.py")).read())
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "combine_pose_and_limb"][0]
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_occlusion_models.py", "exec"), ns)
        ns.update(config=cfg, wandb=wandb_stub, mb=ref["metrics_batch"].Metrics)
        s = base_self(8)
        s.leg_lifter, s.torso_lifter = module("Leg_Lifter", 7, leg), module("Torso_Lifter", 10, torso)
        s.left_lifter, s.right_lifter = module("Left_Right_Lifter", 11, left), module("Left_Right_Lifter", 11, right)
        for i, n in enumerate(OS.OCC_NAMES):
            m = getattr(M, cls[n])(use_batchnorm=False, num_joints=nin[n])
            m.load_state_dict(ON.init_predictor_params(nin[n], nout[n], 100 + i), strict=False)
            setattr(s, attr[n], m)
        s.metrics, s.current_epoch = ref["metrics"].Metrics(), 0
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns["validation_step"](s, val_batch, 0)
        for k, v in s.losses.__dict__.items():
            if k.startswith("pa_") or k.startswith("mpjpe_scaled_"):
                out["occval_" + k] = np.float64(v)
    finally:
        torch.Tensor.cuda = cuda_orig
    np.savez_compressed(os.path.join(OUT, "ref_steps.npz"), **out)
    print("ref_steps.npz written:", {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


def main():
    if "--only-reference-steps" in sys.argv:
        reference_training_steps(import_reference())
        return
    if "--only-metrics-np-extra" in sys.argv:
        metrics_np_extra(import_reference())
        return
    if "--only-script-functions" in sys.argv:
        script_functions()
        return
    if "--only-cli-contract" in sys.argv:
        cli_contract()
        return
    if "--only-module-contract" in sys.argv:
        module_contract(import_reference())
        return
    if "--only-helpers-extra" in sys.argv:        # added later: does not disturb the RNG streams of the other fixtures
        helpers_extra(import_reference())
        return
    from oracle import flow as OF, geometry as OG, metrics as OM, nets as ON, steps as OS
    from links_b200.synth import synth_poses, synth_pred_3d
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---------------- networks (models_def.py) ----------------
    nets = {}
    for cls, nj, seed in (("Leg_Lifter", 7, 11), ("Torso_Lifter", 10, 12), ("Left_Right_Lifter", 11, 13)):
        p = ON.init_lifter_params(nj, seed)
        m = getattr(ref["models_def"], cls)(use_batchnorm=False, num_joints=nj, use_dropout=False, d_rate=0.25)
        missing = m.load_state_dict(p, strict=False)
        assert all(".bn" in k for k in missing.missing_keys) and not missing.unexpected_keys
        x = torch.randn(16, 2 * nj) * 0.15
        x.requires_grad_(True)
        xd, xa = m(x)
        (xd.square().sum() + xa.sum()).backward()
        pr = OS.params_require_grad(p)
        x2 = x.detach().clone().requires_grad_(True)
        od, oa = ON.lifter_forward(x2, pr)
        (od.square().sum() + oa.sum()).backward()
        assert torch.equal(od, xd) and torch.equal(oa, xa), cls
        assert torch.allclose(x2.grad, x.grad, rtol=0, atol=0)
        gw = m.res_pose2.l1.weight.grad
        assert torch.allclose(pr["res_pose2.l1.weight"].grad, gw, rtol=1e-6, atol=1e-9)
        nets[cls + "_seed"] = np.int64(seed)
        nets[cls + "_x"] = t(x)
        nets[cls + "_xd"] = t(xd)
        nets[cls + "_xa"] = t(xa)
        nets[cls + "_dx"] = t(x.grad)
        nets[cls + "_dW_res_pose2_l1_sum"] = np.float64(gw.double().sum().item())
        nets[cls + "_dW_res_pose2_l1_abs"] = np.float64(gw.double().abs().sum().item())
    for cls, nj, od_, seed in (("Occluded_Limb_Predictor", 14, 9, 21), ("Occluded_Legs_Predictor", 11, 18, 22),
                               ("Occluded_Torso_Predictor", 7, 30, 23), ("Occluded_Left_Right_Predictor", 11, 18, 24)):
        p = ON.init_predictor_params(nj, od_, seed)
        m = getattr(ref["models_def"], cls)(use_batchnorm=False, num_joints=nj)
        missing = m.load_state_dict(p, strict=False)
        assert all(".bn" in k for k in missing.missing_keys) and not missing.unexpected_keys
        x = torch.randn(16, 3 * nj)
        y = m(x)
        assert torch.equal(ON.predictor_forward(x, p), y), cls
        nets[cls + "_seed"] = np.int64(seed)
        nets[cls + "_x"] = t(x)
        nets[cls + "_y"] = t(y)
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **nets)

    # ---------------- index maps (helpers.py) -- integer, bit-exact ----------------
    H = ref["helpers"]
    idx = {}
    B = 6
    a34 = torch.arange(B * 34, dtype=torch.float32).reshape(B, 34)
    l, r = H.split_data_left_right(a34)
    ol, or_ = OG.split_data_left_right(a34)
    assert torch.equal(l, ol) and torch.equal(r, or_)
    idx["split_lr_in"], idx["split_lr_left"], idx["split_lr_right"] = t(a34), t(l), t(r)
    a51 = torch.arange(B * 51, dtype=torch.float32).reshape(B, 3, 17)
    l3, r3 = H.split_data_left_right_3d(a51)
    ol3, or3 = OG.split_data_left_right_3d(a51)
    assert torch.equal(l3, ol3) and torch.equal(r3, or3)
    idx["split3d_in"], idx["split3d_left"], idx["split3d_right"] = t(a51), t(l3), t(r3)
    la = torch.arange(B * 11, dtype=torch.float32).reshape(B, 11)
    ra = 1000 + torch.arange(B * 11, dtype=torch.float32).reshape(B, 11)
    for ch in ("left", "right"):
        c = H.combine_left_right_pred_1d(la, ra, ch).reshape(-1, 17)
        assert torch.equal(c, OG.combine_left_right_1d(la, ra, ch))
        idx["combine1d_" + ch] = t(c)
        for dims, fn in ((2, H.combine_left_right_pred_2d), (3, H.combine_left_right_pred_3d)):
            ll = torch.arange(B * dims * 11, dtype=torch.float32).reshape(B, dims * 11)
            rr = 5000 + ll
            c = fn(ll, rr, ch)
            assert torch.equal(c, OG.combine_left_right_nd(ll, rr, ch, dims))
            idx["combine%dd_%s" % (dims, ch)] = t(c)
    idx["combine_la"], idx["combine_ra"] = t(la), t(ra)
    # occlusion gathers (train_occlusion_models.py:176-191 restated in oracle.steps) on the arange pose
    tg, inp = OS.occ_targets_inputs(a51)
    for n in OS.OCC_NAMES:
        idx["occ_target_" + n] = t(tg[n])
        idx["occ_input_" + n] = t(inp[n])
    np.savez_compressed(os.path.join(OUT, "index_maps.npz"), **idx)

    # ---------------- geometry (helpers.py, rotation_conversions.py) ----------------
    geo = {}
    ang = torch.randn(32, 3)
    for conv in ("XYZ", "ZYX", "YXZ", "XZY"):
        Rr = ref["rotation_conversions"].euler_angles_to_matrix(ang, conv)
        assert torch.equal(Rr, OG.euler_angles_to_matrix(ang, conv)), conv
        geo["euler_" + conv] = t(Rr)
    geo["euler_in"] = t(ang)
    p3 = torch.randn(32, 51)
    p3[:, 34:] = p3[:, 34:].abs() + 5.0
    pp = H.perspective_projection(p3)
    assert torch.equal(pp, OG.perspective_projection(p3))
    geo["proj_in"], geo["proj_out"] = t(p3), t(pp)
    bl = H.get_bone_lengths_all(p3)
    assert torch.equal(bl, OG.get_bone_lengths_all(p3))
    geo["bones_out"] = t(bl)
    noise_in = torch.randn(8, 34)
    torch.manual_seed(5)
    ref_noisy = H.add_noise(noise_in, 0.2)
    torch.manual_seed(5)
    nz = torch.randn_like(noise_in)
    assert torch.equal(ref_noisy, OG.add_noise(noise_in, nz, 0.2))
    raw2d = np.random.RandomState(3).normal(size=(8, 34)) * 100 + 500
    nh = H.normalize_head(raw2d.copy())
    assert np.array_equal(nh, OG.normalize_head(raw2d.copy()))
    nht = H.normalize_head_test(raw2d.copy())
    assert np.array_equal(nht, OG.normalize_head_test(raw2d.copy()))
    geo["normhead_in"], geo["normhead_out"], geo["normhead_test_out"] = raw2d, nh, nht
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **geo)

    # ---------------- metrics ----------------
    met = {}
    p2d, gt = synth_poses(96, seed=7)
    pred = synth_pred_3d(gt, seed=8, noise_mm=40.0, scale=0.0125, mirror_frac=0.25)
    gt_t, pred_t = torch.from_numpy(gt), torch.from_numpy(pred)
    mb = ref["metrics_batch"].Metrics()
    met["gt"], met["pred"] = gt, pred
    for nj, rj in ((17, 0), (16, 6)):
        g, p = gt_t[:, :3 * 17].reshape(-1, 3, 17)[:, :, :nj].reshape(-1, 3 * nj), \
            pred_t.reshape(-1, 3, 17)[:, :, :nj].reshape(-1, 3 * nj)
        kw = dict(num_joints=nj, root_joint=rj)
        for scaling in (True, False):
            e = mb.mpjpe(g, p, use_scaling=scaling, **kw)
            assert torch.equal(e, OM.mpjpe(g, p, use_scaling=scaling, **kw))
            met["mpjpe_j%d_s%d" % (nj, scaling)] = t(e)
        e = mb.PCK(g, p, **kw)
        assert torch.equal(e, OM.pck(g, p, **kw))
        met["pck_j%d" % nj] = t(e)
        e = mb.AUC(g, p, **kw)
        assert torch.allclose(e, OM.auc(g, p, **kw), rtol=0, atol=0)
        met["auc_j%d" % nj] = t(e)
        ga = mb.get_all(g, p, **kw)
        oa = OM.get_all(g, p, **kw)
        for k in ga:
            assert torch.allclose(ga[k], oa[k], rtol=0, atol=0), k
            met["getall_%s_j%d" % (k, nj)] = t(ga[k])
        e = mb.pmpjpe(g, p, num_joints=nj)
        o = OM.pmpjpe_batch(g, p, num_joints=nj)
        assert torch.allclose(e, o, rtol=1e-4, atol=1e-3), (e - o).abs().max()  # torch.svd vs linalg.svd
        met["pmpjpe_batch_j%d" % nj] = t(e)
    mn = ref["metrics"].Metrics()
    gt64, pr64 = gt.astype(np.float64), pred.astype(np.float64)
    for refl in ("best", True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e = np.array([mn.pmpjpe(gt64[i].reshape(-1, 51), pr64[i].reshape(-1, 51), reflection=refl)
                          for i in range(gt.shape[0])])
        o = np.array([OM.pmpjpe_best_np(gt64[i].reshape(-1, 51), pr64[i].reshape(-1, 51), reflection=refl)
                      for i in range(gt.shape[0])])
        assert np.allclose(e, o, rtol=0, atol=1e-9), refl
        met["pmpjpe_np_%s" % refl] = e
    ob = OM.pmpjpe_best_batch(gt64, pr64)
    assert np.allclose(ob, met["pmpjpe_np_best"], rtol=0, atol=1e-8)
    # plain numpy mpjpe (metrics.py:8-33)
    e = np.array([mn.mpjpe(gt64[i].reshape(-1, 51), pr64[i].reshape(-1, 51)) for i in range(8)])
    met["mpjpe_np"] = e
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **met)

    # ---------------- steps (oracle-composed from the pinned pieces; flow unpinned) ----------------
    st = {}
    B = 8
    x2d, _ = synth_poses(B, seed=31)
    x = torch.from_numpy(x2d)
    g = torch.Generator().manual_seed(77)
    noise = torch.randn(B, 34, generator=g)
    eps_x = torch.randn(2 * B, generator=g)
    u_y = torch.rand(2 * B, generator=g)
    full_flow = OF.init_flow_params(34, 40, perturb=0.3)
    u = OS.sample_poses(x, full_flow, noise)
    st["x"], st["noise"], st["eps_x"], st["u_y"], st["u"] = t(x), t(noise), t(eps_x), t(u_y), t(u)
    leg, torso = OS.params_require_grad(ON.init_lifter_params(7, 11)), OS.params_require_grad(ON.init_lifter_params(10, 12))
    lf, tf = OF.init_flow_params(14, 41, perturb=0.3), OF.init_flow_params(20, 42, perturb=0.3)
    aux = {}
    out = OS.lt_step(u, leg, torso, lf, tf, eps_x, u_y, aux=aux)
    out["loss"].backward()
    for k, v in out.items():
        st["lt_" + k] = np.float64(v.item())
    st["lt_rot_2d"], st["lt_pred"] = t(aux["rot_2d"]), t(aux["pred"])
    st["lt_dW_leg_upscale"] = t(leg["upscale.weight"].grad)
    st["lt_dW_torso_angles"] = t(torso["angles.weight"].grad)
    left, right = OS.params_require_grad(ON.init_lifter_params(11, 13)), OS.params_require_grad(ON.init_lifter_params(11, 14))
    lff, rff = OF.init_flow_params(22, 43, perturb=0.3), OF.init_flow_params(22, 44, perturb=0.3)
    aux = {}
    out = OS.lr_step(u, left, right, lff, rff, eps_x, u_y, aux=aux)
    out["loss"].backward()
    for k, v in out.items():
        st["lr_" + k] = np.float64(v.item())
    st["lr_rot_2d_left"], st["lr_rot_2d_right"] = t(aux["rot_2d_left"]), t(aux["rot_2d_right"])
    st["lr_dW_left_upscale"] = t(left["upscale.weight"].grad)
    # flow step (config #1), fp64 twin for the known-answer log-det check is in tests
    fp = OS.params_require_grad(OF.init_flow_params(34, 40, perturb=0.3))
    out = OS.flow_step(x, fp, noise)
    out["loss"].backward()
    for k, v in out.items():
        st["flow_" + k] = np.float64(v.item())
    st["flow_dW_b0_l0"] = t(fp["module_list.0.subnet.0.weight"].grad)
    z, ld = OF.inn_forward(x, OF.init_flow_params(34, 40, perturb=0.3))
    st["flow_z"], st["flow_ld"] = t(z), t(ld)
    np.savez_compressed(os.path.join(OUT, "steps.npz"), **st)
    helpers_extra(ref)
    module_contract(ref)
    cli_contract()
    script_functions()
    metrics_np_extra(ref)
    reference_training_steps(ref)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-20s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
