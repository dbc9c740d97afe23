python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -12 > gpurun_out/r2_t5.log
for i in 1 2; do SGMCMC_B200_LIB=$PWD/build/variants/lib_oldksd.so python tools/ksd_bench.py 2>&1 | grep "^tc" | sed 's/^/old /' >> gpurun_out/r2_ksd_ab.log; python tools/ksd_bench.py 2>&1 | grep "^tc" | sed 's/^/new /' >> gpurun_out/r2_ksd_ab.log; done
ncu --set full --clock-control none --import-source on -k regex:ksd_tc_kernel -c 1 -o gpurun_out/r2_ksd_aug python tools/ksd_bench.py > gpurun_out/r2_ksd_ncu.log 2>&1
python tools/phase_timing.py mlp_psgld > gpurun_out/r2_mlp_phases.log 2>&1
compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_cases.py > gpurun_out/r2_memcheck.log 2>&1
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_cases.py vec cluster pmf opt > gpurun_out/r2_racecheck.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
cat gpurun_out/r2_t5.log gpurun_out/r2_ksd_ab.log gpurun_out/r2_mlp_phases.log; tail -5 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log; tail -c 600 gpurun_out/r2_bench3.json; tail -3 gpurun_out/r2_bench3.err
