#!/bin/bash
# same-box A/B of the pointwise GEMM producers: register loads vs the TMA-fed raw ring (needs the -DCFNET_AB build)
export CFNET_LIB=$PWD/coarse_fine_networks_b200/libcfnet_b200_ab.so
for rep in 1 2; do
  echo "=== TMA=0 (register-load producers) rep $rep"; CFNET_P2_TMA=0 python tools/bench_pw.py "$@" 2>&1 | grep -v "^weight\|wgrad"
  echo "=== TMA=1 (TMA-fed raw ring) rep $rep"; CFNET_P2_TMA=1 python tools/bench_pw.py "$@" 2>&1 | grep -v "^weight\|wgrad"
done
