timeout 1500 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_randint_ilp.log
for i in 1 2; do
echo "== new" >> gpurun_out/r2_randint_ilp.log; python tools/prof_c1.py 1000 2>&1 | tail -1 >> gpurun_out/r2_randint_ilp.log; python tools/prof_c2.py 2>&1 | tail -1 >> gpurun_out/r2_randint_ilp.log
echo "== prev" >> gpurun_out/r2_randint_ilp.log; SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so python tools/prof_c1.py 1000 2>&1 | tail -1 >> gpurun_out/r2_randint_ilp.log; SGMCMC_B200_LIB=$PWD/build/variants/lib_prev.so python tools/prof_c2.py 2>&1 | tail -1 >> gpurun_out/r2_randint_ilp.log
done
cat gpurun_out/r2_randint_ilp.log
