for i in 1 2; do
for v in "" mlpu2 mlpu6 mlpu8; do
  echo "== variant ${v:-default(U=4)}" >> gpurun_out/r2_mlp_unroll.log
  if [ -z "$v" ]; then python tools/prof_big.py mlp_psgld 256 16 3 2>&1 | tail -1 >> gpurun_out/r2_mlp_unroll.log; else SGMCMC_B200_LIB=$PWD/build/variants/lib_$v.so python tools/prof_big.py mlp_psgld 256 16 3 2>&1 | tail -1 >> gpurun_out/r2_mlp_unroll.log; fi
done; done
cat gpurun_out/r2_mlp_unroll.log
