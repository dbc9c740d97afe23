python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fullbatch or cv_row or ksd_pipeline or posterior" 2>&1 | tail -5 > gpurun_out/r2_t15.log
for mb in 1024 256 96 64 48 32 16; do echo "== W budget $mb MB" >> gpurun_out/r2_k2_blocking.log; SGMCMC_K2_W_MB=$mb python tools/k2_bench.py 2>&1 | grep "^K2" | tail -2 >> gpurun_out/r2_k2_blocking.log; done
cat gpurun_out/r2_t15.log gpurun_out/r2_k2_blocking.log
