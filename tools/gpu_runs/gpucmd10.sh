M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__cluster_dim_x,launch__grid_size,launch__block_size"
# C2, 1024 chains (launch 0 = init, 1.. = run): full set on the third launch
ncu --set full --clock-control none --import-source on -k regex:vec_sampler_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_c2_sgld python tools/prof_c2.py > gpurun_out/r2_ncu_c2.log 2>&1
# C2 with 128 chains (the 8-GPU shard) and C1 (cluster per chain): metric list
ncu --metrics $M --clock-control none -k regex:vec_sampler_kernel --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2_c2_128_metrics.csv python tools/cluster_sweep.py c2 128 10 "0x0" > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:vec_sampler_kernel --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2_c1_metrics.csv python tools/cluster_sweep.py c1 "0x0" > /dev/null 2>&1
# MLP and PMF samplers, K2 scores / accumulate
ncu --metrics $M --clock-control none -k regex:big_sampler_kernel --launch-skip 1 -c 1 --csv --log-file gpurun_out/r2_mlp_metrics.csv python tools/prof_big.py mlp_psgld 148 4 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:big_sampler_kernel --launch-skip 1 -c 1 --csv --log-file gpurun_out/r2_pmf_metrics.csv python tools/prof_big.py pmf_sgnht 512 16 2 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k2_ --launch-skip 20 -c 2 --csv --log-file gpurun_out/r2_k2_metrics.csv python tools/k2_bench.py > /dev/null 2>&1
ls -la gpurun_out/r2_c2_sgld.ncu-rep; tail -3 gpurun_out/r2_ncu_c2.log; for f in c2_128 c1 mlp pmf k2; do echo == $f; tail -n +1 gpurun_out/r2_${f}_metrics.csv | grep -v "^==" | cut -d, -f5,13- | head -40; done
