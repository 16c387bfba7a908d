set -x
python -m pytest tests/test_kernels_geom.py tests/test_gpu_mlp_steps.py tests/test_gpu_parity_baseline_sizes.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_t_geom.log
python bench_kernels.py > gpurun_out/r02_kernels_isolated_v2.json 2> gpurun_out/r02_bk2.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_write.sum"
ncu --metrics $M --clock-control none -k regex:"geom_|eval_lift" --csv --log-file gpurun_out/r02_ncu_geom_v2.csv python scratch/geom_prof.py lt > gpurun_out/r02_ncu_geom_v2.log 2>&1
