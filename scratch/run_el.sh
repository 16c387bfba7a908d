set -x
for v in el0 el1; do
  export LINKS_B200_LIB=$PWD/scratch/variants/$v/liblinks_b200.so
  echo "== $v" >> gpurun_out/r02_evict_last_ab.txt
  CHAIN_PROF_ONLY="bwd0+wgrad,bwd0+wgrad+adam,wgrad" python scratch/chain_prof.py 1024 >> gpurun_out/r02_evict_last_ab.txt 2>&1
  CHAIN_PROF_ONLY="bwd0+wgrad+adam" python scratch/chain_prof.py 8192 >> gpurun_out/r02_evict_last_ab.txt 2>&1
  python scratch/sweep_reserve.py 1024 -1 >> gpurun_out/r02_evict_last_ab.txt 2>&1
  python scratch/sweep_reserve.py 8192 24 >> gpurun_out/r02_evict_last_ab.txt 2>&1
done
unset LINKS_B200_LIB
