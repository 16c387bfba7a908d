set -x
python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_t_dp_final.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
