python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py -m gpu -q --maxfail=8 2>&1 | tail -12 > gpurun_out/r2_t13.log
python tools/cluster_sweep.py c1 "0x0,16x8,8x8,16x4" > gpurun_out/r2_sweep13.log 2>&1
SGMCMC_NO_LOOKAHEAD=1 python tools/cluster_sweep.py c1 "0x0" >> gpurun_out/r2_sweep13.log 2>&1
python tools/cluster_sweep.py c2 1024 10 "0x0" >> gpurun_out/r2_sweep13.log 2>&1
python tools/cluster_sweep.py c2 128 10 "0x0" >> gpurun_out/r2_sweep13.log 2>&1
python tools/cluster_sweep.py c2 16 10 "0x0" >> gpurun_out/r2_sweep13.log 2>&1
SGMCMC_NO_LOOKAHEAD=1 python tools/cluster_sweep.py c2 16 10 "0x0" >> gpurun_out/r2_sweep13.log 2>&1
cat gpurun_out/r2_t13.log gpurun_out/r2_sweep13.log
