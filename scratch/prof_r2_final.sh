# Final evidence of round 2 (one B200): GPU tests, smoke, bench, reference arm, ncu launch list, ncu --set full of the chain launches
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_final.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err
python bench_kernels.py > gpurun_out/r02_kernels_isolated_final.json 2> gpurun_out/r02_bk_final.err
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra --no-graph > gpurun_out/r02_plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra --no-graph > gpurun_out/r02_ncu_launches.log 2>&1
python scratch/chain_ncu.py 1024 pure > gpurun_out/r02_chain_b1024_pure_times.txt 2>&1 && \
ncu --set full --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv --page raw --log-file gpurun_out/r02_chains_b1024_pure_raw.csv python scratch/chain_ncu.py 1024 pure > gpurun_out/r02_chain_b1024_pure_ncu.log 2>&1
ls -la gpurun_out | tail -12
