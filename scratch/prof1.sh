set -x
python bench.py --steps 3 --warmup 3 --skip-cpu --no-graph > gpurun_out/plain_r1b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --skip-cpu --no-graph > gpurun_out/ncu_r1b.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
