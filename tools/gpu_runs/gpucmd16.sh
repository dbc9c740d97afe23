python -m pytest tests/test_gpu_big_families.py tests/test_gpu_baseline_geometries.py tests/test_optimizer.py -m gpu -q --maxfail=6 2>&1 | tail -6 > gpurun_out/r2_t16.log
for lib in prev new; do echo "== $lib" >> gpurun_out/r2_pair.log; for w in "mlp_psgld 256 16 3" "mlp_sghmc 256 4 3" "pmf_sgnht 512 16 3" "pmf_badodabcv 512 16 3"; do if [ $lib = prev ]; then export SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so; else unset SGMCMC_B200_LIB; fi; python tools/prof_big.py $w 2>&1 | tail -1 >> gpurun_out/r2_pair.log; done; done
cat gpurun_out/r2_t16.log gpurun_out/r2_pair.log
