python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph > gpurun_out/plain_r1d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph > gpurun_out/ncu_r1d.log 2>&1
