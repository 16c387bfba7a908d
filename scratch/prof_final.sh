set -x
python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/ncu_final1.log 2>&1
ncu --set full --clock-control none -k regex:gemm_grouped -s 320 -c 12 --csv --page raw --log-file gpurun_out/gemm_final_raw.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/ncu_final2.log 2>&1
ncu --set full --clock-control none -k regex:flow_tc -s 20 -c 4 --csv --page raw --log-file gpurun_out/flow_final_raw.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --serial > gpurun_out/ncu_final3.log 2>&1
ls -la gpurun_out
