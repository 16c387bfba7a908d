set -x
python -m pytest tests/test_kernels_geom.py tests/test_gpu_mlp_steps.py tests/test_gpu_parity_baseline_sizes.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_t_geom.log
python bench_kernels.py > gpurun_out/r02_kernels_isolated_v14.json 2> gpurun_out/r02_bk14.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
ncu --metrics $M --clock-control none -k regex:"geom_" --csv --log-file gpurun_out/r02_ncu_geom_v14.csv python scratch/geom_prof.py lt > gpurun_out/r02_ncu_geom_v14.log 2>&1
