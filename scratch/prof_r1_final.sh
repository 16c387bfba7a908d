# final profile pass of round 1 (run under gpurun; every ncu command follows a clean plain run of the same program)
set -x
timeout 120 python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/plain_r1f.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/ncu_r1f_1.log 2>&1
timeout 240 ncu --set full --clock-control none -k regex:gemm_grouped -s 320 -c 12 --csv --page raw --log-file gpurun_out/gemm_r1f_raw.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/ncu_r1f_2.log 2>&1
ls -la gpurun_out | tail -5
