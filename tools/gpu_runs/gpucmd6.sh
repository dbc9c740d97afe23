python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -8 > gpurun_out/r2_t6.log
python tools/k2_bench.py > gpurun_out/r2_k2.log 2>&1
python tools/prof_big.py mlp_psgld 148 4 3 > gpurun_out/r2_mlp.log 2>&1
python tools/prof_big.py mlp_sghmc 256 4 3 >> gpurun_out/r2_mlp.log 2>&1
python tools/ksd_bench.py 2>&1 | grep "^tc" >> gpurun_out/r2_k2.log
cat gpurun_out/r2_t6.log gpurun_out/r2_k2.log gpurun_out/r2_mlp.log
