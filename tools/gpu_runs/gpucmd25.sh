timeout 1200 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_c1_async.log
for i in 1 2; do
echo "== new" >> gpurun_out/r2_c1_async.log; python tools/prof_c1.py 1000 2>&1 | tail -1 >> gpurun_out/r2_c1_async.log
echo "== prev" >> gpurun_out/r2_c1_async.log; SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so python tools/prof_c1.py 1000 2>&1 | tail -1 >> gpurun_out/r2_c1_async.log
done
python tools/cluster_sweep.py c2 2>&1 | tail -6 >> gpurun_out/r2_c1_async.log
cat gpurun_out/r2_c1_async.log
