python -m pytest tests/test_gpu_big_families.py tests/test_gpu_baseline_geometries.py tests/test_optimizer.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_mlp_issuer_fence.log

for i in 1 2; do
for w in "mlp_psgld 256 16 3" "mlp_sghmc 256 4 3"; do
  echo "== new $w" >> gpurun_out/r2_mlp_issuer_fence.log; python tools/prof_big.py $w 2>&1 | tail -1 >> gpurun_out/r2_mlp_issuer_fence.log
  echo "== prev $w" >> gpurun_out/r2_mlp_issuer_fence.log; SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so python tools/prof_big.py $w 2>&1 | tail -1 >> gpurun_out/r2_mlp_issuer_fence.log
done; done
cat gpurun_out/r2_mlp_issuer_fence.log
