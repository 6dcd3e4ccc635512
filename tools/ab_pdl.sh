#!/bin/bash
# Same-box A/B of programmatic dependent launch (AB build: CFNET_PDL=0 = plain launches, 1 = programmatic edges, early trigger
# in the table kernels only; "pdlwait" variant = edges, no early trigger anywhere) and of the weight-gradient side stream.
P=$PWD/coarse_fine_networks_b200
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['parity']['rel_linf'], d['config'].get('cuda_graph'))"; }
B="timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-eager-baseline"
for i in 1 2; do
  CFNET_LIB=$P/libcfnet_b200_ab.so CFNET_PDL=0 CF_WGRAD_STREAM_ROWS=0 $B 2>/dev/null | line "pdl=0 wgrad_rows=0"
  CFNET_LIB=$P/libcfnet_b200_pdlwait.so CF_WGRAD_STREAM_ROWS=0 $B 2>/dev/null | line "edge-only wgrad_rows=0"
  CFNET_LIB=$P/libcfnet_b200_ab.so CFNET_PDL=1 CF_WGRAD_STREAM_ROWS=0 $B 2>/dev/null | line "edge+table-trigger wgrad_rows=0"
  CFNET_LIB=$P/libcfnet_b200_ab.so CFNET_PDL=1 CF_WGRAD_STREAM_ROWS=65536 $B 2>/dev/null | line "edge+table-trigger wgrad_rows=65536"
  CFNET_LIB=$P/libcfnet_b200_ab.so CFNET_PDL=1 CF_WGRAD_STREAM_ROWS=262144 $B 2>/dev/null | line "edge+table-trigger wgrad_rows=262144"
done
