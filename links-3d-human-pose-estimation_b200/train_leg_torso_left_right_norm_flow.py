"""Drop-in for the reference's train_leg_torso_left_right_norm_flow.py: same CLI flag (:27-31), same step (NLL of
the leg / torso / left / right parts of the data and of samples drawn from the frozen full-pose flow, four
Adam(2e-4, wd 1e-5) optimisers, ExponentialLR 0.95 per epoch) and the same checkpoint names
(mpi_norm_flow_{left,right,legs,torso}_2.pt, saved every epoch, FrEIA key layout; :195-198)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from links_b200 import init as INIT  # noqa: E402
from links_b200.flowtrain import PartFlowTrainer  # noqa: E402
from links_b200.harness import GAMMA, LR0, add_common_args, ckpt_paths, dist_setup, load_state, make_loader  # noqa: E402

parser = argparse.ArgumentParser(description='Train 2D INN')
parser.add_argument("-l", "--left_right_side_keypoints", help="number of key-points in each split", type=int, default=22)
add_common_args(parser, batch=256, epochs=100)

FILES = {"left": "mpi_norm_flow_left_2.pt", "right": "mpi_norm_flow_right_2.pt", "legs": "mpi_norm_flow_legs_2.pt",
         "torso": "mpi_norm_flow_torso_2.pt"}
WIDTH = {"legs": 14, "torso": 20, "left": 22, "right": 22}

if __name__ == "__main__":
    args = parser.parse_args()
    if args.left_right_side_keypoints != 22:
        raise NotImplementedError("split_data_left_right (utils/helpers.py:55-65) yields 11 joints = 22 values per side")
    rank, world, pg = dist_setup()
    full = load_state(ckpt_paths(args.weights_dir, "full_flow_parts"), lambda: INIT.init_flow_params(34, 40), args.random_init)
    parts = {n: INIT.init_flow_params(WIDTH[n], 50 + i + args.seed, perturb=0.0) for i, n in enumerate(PartFlowTrainer.NAMES)}
    loader = make_loader(args, rank, world)
    trainer = PartFlowTrainer(full, parts, loader.batch, lr=LR0, weight_decay=1e-5, process_group=pg)
    gen_dev = torch.Generator(device="cuda").manual_seed(args.seed * 7919 + rank)
    n_steps, lr, t0 = 0, LR0, time.time()
    done = False
    for epoch in range(args.epochs):
        trainer.set_lr(lr)
        for xb in loader:
            trainer.x.copy_(xb, non_blocking=True)
            trainer.noise.normal_(generator=gen_dev)
            trainer.run(use_graph=not args.no_graph)   # CUDA-graph replay from the third step on
            n_steps += 1
            if rank == 0 and n_steps % args.log_every == 0:
                d = trainer.loss_dict()
                print("epoch %d step %d  %s  (%.0f poses/s)" % (epoch, n_steps, " ".join("%s=%.5f" % kv for kv in d.items()),
                                                                n_steps * args.batch / (time.time() - t0)), flush=True)
            if args.steps and n_steps >= args.steps:
                done = True
                break
        lr *= GAMMA
        if rank == 0 and not args.no_save:
            os.makedirs(args.weights_dir, exist_ok=True)
            for n, f in FILES.items():
                torch.save({k: v.detach().cpu().clone() for k, v in trainer.state_dict(n).items()},
                           os.path.join(args.weights_dir, f))
        if done:
            break
