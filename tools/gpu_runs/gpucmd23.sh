for i in 1 2; do
for v in "" pmf3a pmf3b; do
  echo "== variant ${v:-default (2 CTAs x 256 x 128 regs, U=4)}" >> gpurun_out/r2_pmf_3cta.log
  if [ -z "$v" ]; then python tools/prof_big.py pmf_sgnht 512 16 3 2>&1 | tail -1 >> gpurun_out/r2_pmf_3cta.log; else SGMCMC_B200_LIB=$PWD/build/variants/lib_$v.so python tools/prof_big.py pmf_sgnht 512 16 3 2>&1 | tail -1 >> gpurun_out/r2_pmf_3cta.log; fi
done; done
cat gpurun_out/r2_pmf_3cta.log
