#!/bin/bash
# ncu --set full captures of the round-2 hot kernels (one launch each, after warm-up), reports into gpurun_out/
mkdir -p gpurun_out
for k in pw_tc pw_swish pw_dgrad3 wgrad1 wgrad3 dw_fused; do
  case $k in
    pw_*) pat="pw_tc2_kernel";;
    wgrad*) pat="pw_wgrad_tc_kernel";;
    dw_fused) pat="dw3_kernel";;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 4 -c 1 -f -o gpurun_out/r02_full_$k \
      python profiles/run_kernel.py $k 4 64 > gpurun_out/r02_ncu_$k.log 2>&1
  tail -1 gpurun_out/r02_ncu_$k.log
done
