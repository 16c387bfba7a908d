python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph > gpurun_out/plain_r1c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 800 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph > gpurun_out/ncu_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -s 300 -c 6 -o gpurun_out/gemm_v3 -f python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph > gpurun_out/ncu_r1c2.log 2>&1
