for w in "mlp_psgld 256 16 3" "mlp_sghmc 256 4 3"; do python tools/prof_big.py $w 2>&1 | tail -1 >> gpurun_out/r2_pmf_threads.log; done
for t in 128 192 256 384 512; do echo "== PMF threads $t" >> gpurun_out/r2_pmf_threads.log; SGMCMC_PMF_THREADS=$t python tools/prof_big.py pmf_sgnht 512 16 3 2>&1 | tail -1 >> gpurun_out/r2_pmf_threads.log; SGMCMC_PMF_THREADS=$t python tools/prof_big.py pmf_badodabcv 512 16 3 2>&1 | tail -1 >> gpurun_out/r2_pmf_threads.log; done
cat gpurun_out/r2_pmf_threads.log
