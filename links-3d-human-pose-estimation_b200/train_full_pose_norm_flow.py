"""Drop-in for the reference's train_full_pose_norm_flow.py: same CLI flag (-n/--num_keypoints, :21-25), same step
(NLL of the data + NLL of the flow's own noisy samples, Adam 2e-4 wd 1e-5, ExponentialLR 0.95 per epoch) and the same
checkpoint name (models/norm_flow_sampling.pt, saved every epoch, FrEIA key layout)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from links_b200 import init as INIT  # noqa: E402
from links_b200.flowtrain import FlowTrainStep  # noqa: E402
from links_b200.harness import GAMMA, LR0, add_common_args, dist_setup, make_loader  # noqa: E402

parser = argparse.ArgumentParser(description='Train 2D INN')
parser.add_argument("-n", "--num_keypoints", help="number of keypoints", type=int, default=34)
add_common_args(parser, batch=4 * 64, epochs=100)

if __name__ == "__main__":
    args = parser.parse_args()
    if args.num_keypoints != 34:
        raise NotImplementedError("the sampling block zeroes the root joint of a 17-joint pose (reference :84-86)")
    rank, world, pg = dist_setup()
    params = INIT.init_flow_params(34, 40 + args.seed, perturb=0.0)
    loader = make_loader(args, rank, world)
    step = FlowTrainStep(34, params, loader.batch, lr=LR0, weight_decay=1e-5, process_group=pg)
    gen_dev = torch.Generator(device="cuda").manual_seed(args.seed * 7919 + rank)
    n_steps, lr, t0 = 0, LR0, time.time()
    done = False
    for epoch in range(args.epochs):
        step.set_lr(lr)
        for xb in loader:
            step.x.copy_(xb, non_blocking=True)
            step.noise.normal_(generator=gen_dev)
            step.run(use_graph=not args.no_graph)      # CUDA-graph replay from the third step on
            n_steps += 1
            if rank == 0 and n_steps % args.log_every == 0:
                print("epoch %d step %d  loss=%.5f  (%.0f poses/s)" % (epoch, n_steps, step.loss_dict()["loss"],
                                                                      n_steps * args.batch / (time.time() - t0)), flush=True)
            if args.steps and n_steps >= args.steps:
                done = True
                break
        lr *= GAMMA
        if rank == 0 and not args.no_save:
            os.makedirs(args.weights_dir, exist_ok=True)
            torch.save({k: v.detach().cpu().clone() for k, v in step.state_dict().items()},
                       os.path.join(args.weights_dir, "norm_flow_sampling.pt"))
        if done:
            break
