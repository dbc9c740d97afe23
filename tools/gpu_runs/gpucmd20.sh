timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_geometries.py -x -q -m gpu -k "ksd or KSD" 2>&1 | tail -5 > gpurun_out/r2_ksd_merged.log
for i in 1 2; do
echo "== cluster 2" >> gpurun_out/r2_ksd_merged.log; timeout 300 python tools/ksd_bench.py 2>&1 | grep "^tc" >> gpurun_out/r2_ksd_merged.log
echo "== cluster 1 (one CTA per tile)" >> gpurun_out/r2_ksd_merged.log; SGMCMC_KSD_CLUSTER=1 timeout 300 python tools/ksd_bench.py 2>&1 | grep "^tc" >> gpurun_out/r2_ksd_merged.log
done
cat gpurun_out/r2_ksd_merged.log
