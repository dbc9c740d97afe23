L=$PWD/build/variants
for v in mlpbase mlpcg mlpiss mlpisscg; do
  echo "== $v" >> gpurun_out/r2_mlp_ab.log
  SGMCMC_B200_LIB=$L/lib_$v.so timeout 300 python tools/prof_big.py mlp_psgld 148 4 3 >> gpurun_out/r2_mlp_ab.log 2>&1
  SGMCMC_B200_LIB=$L/lib_$v.so timeout 300 python tools/prof_big.py mlp_sghmc 256 4 3 >> gpurun_out/r2_mlp_ab.log 2>&1
done
SGMCMC_B200_LIB=$L/lib_mlpiss.so timeout 900 python -m pytest tests/test_gpu_big_families.py tests/test_gpu_baseline_geometries.py -q -k "mlp or c3" --maxfail=5 2>&1 | tail -8 > gpurun_out/r2_mlpiss_tests.log
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
cat gpurun_out/r2_mlp_ab.log gpurun_out/r2_mlpiss_tests.log; tail -2 gpurun_out/r2_smoke.log
