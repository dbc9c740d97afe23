python -m pytest tests/test_gpu_big_families.py tests/test_gpu_baseline_geometries.py tests/test_optimizer.py -m gpu -q --maxfail=6 2>&1 | tail -6 > gpurun_out/r2_t12.log
for q in 0 1; do echo "== SGMCMC_BIG_QUEUE=$q" >> gpurun_out/r2_queue2.log; for w in "mlp_psgld 256 16 3" "mlp_sghmc 256 4 3" "pmf_sgnht 512 16 3" "pmf_badodabcv 512 16 3"; do SGMCMC_BIG_QUEUE=$q python tools/prof_big.py $w 2>&1 | tail -1 >> gpurun_out/r2_queue2.log; done; done
for b in 2 8; do echo "== blocks $b" >> gpurun_out/r2_queue2.log; SGMCMC_BIG_BLOCKS=$b python tools/prof_big.py mlp_psgld 256 16 3 2>&1 | tail -1 >> gpurun_out/r2_queue2.log; SGMCMC_BIG_BLOCKS=$b python tools/prof_big.py pmf_sgnht 512 16 3 2>&1 | tail -1 >> gpurun_out/r2_queue2.log; done
cat gpurun_out/r2_t12.log gpurun_out/r2_queue2.log
