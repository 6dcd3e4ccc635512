#!/bin/bash
# same-box A/B of the pointwise weight-gradient producers: register loads vs the TMA-fed raw ring (-DCFNET_AB build)
export CFNET_LIB=$PWD/coarse_fine_networks_b200/libcfnet_b200_ab.so
for rep in 1 2; do
  echo "=== WG_TMA=0 rep $rep"; CFNET_WG_TMA=0 python tools/bench_pw.py wgrad 2>&1 | grep "wgrad dy"
  echo "=== WG_TMA=1 rep $rep"; CFNET_WG_TMA=1 python tools/bench_pw.py wgrad 2>&1 | grep "wgrad dy"
done
