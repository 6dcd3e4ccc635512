#!/bin/bash
# ncu --set full captures of the round-2 hot kernels (one launch each, after warm-up).  The reports with source are 25 MB
# each: they are summarised ON the box (profiles/summarize.py full + mix, the SASS source page gzipped) and only the
# summaries travel back in gpurun_out/ (64 MiB limit).
mkdir -p gpurun_out
for k in ${@:-pw_tc pw_swish pw_dgrad3 pw_dgrad1 wgrad1 wgrad3 dw_fused}; do
  case $k in
    pw_*) pat="pw_tc2_kernel";;
    wgrad*) pat="pw_wgrad_tc_kernel";;
    dw_fused) pat="dw3_kernel";;
  esac
  rep=/tmp/r02g_full_$k
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 4 -c 1 -f -o $rep \
      python profiles/run_kernel.py $k 4 64 > gpurun_out/r02g_ncu_$k.log 2>&1
  tail -1 gpurun_out/r02g_ncu_$k.log
  python profiles/summarize.py full $rep.ncu-rep gpurun_out/r02g_full_$k.md
  python profiles/summarize.py mix $rep.ncu-rep gpurun_out/r02g_full_$k.md
  ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-12 | gzip -9 > gpurun_out/r02g_src_$k.csv.gz
  rm -f $rep.ncu-rep
done
ls -la gpurun_out | tail -20
