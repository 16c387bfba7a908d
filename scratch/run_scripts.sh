set -x
cd links-3d-human-pose-estimation_b200
python train_leg_torso_lifter.py --synthetic 262144 --batch 1024 --epochs 1 --random-init --no-save --log-every 64 > ../gpurun_out/r02_script_lt_b1024.log 2>&1
python train_left_right_lifter.py --synthetic 262144 --batch 1024 --epochs 1 --random-init --no-save --log-every 64 > ../gpurun_out/r02_script_lr_b1024.log 2>&1
