timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_geometries.py -x -q -m gpu -k "k2 or K2 or fullbatch or grad" 2>&1 | tail -3 > gpurun_out/r2_k2_sel.log
for i in 1 2; do
echo "== new" >> gpurun_out/r2_k2_sel.log; timeout 600 python tools/k2_bench.py 2>&1 | grep "^K2\|max rel" >> gpurun_out/r2_k2_sel.log
echo "== prev" >> gpurun_out/r2_k2_sel.log; SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so timeout 600 python tools/k2_bench.py 2>&1 | grep "^K2\|max rel" >> gpurun_out/r2_k2_sel.log
done
cat gpurun_out/r2_k2_sel.log
