#!/bin/bash
# End-of-round measurement set on one box: GPU suite, smoke, the bench lines of the three workloads, the reference arm, and
# the ncu launch list of the default bench command (shares only).  Outputs in gpurun_out/r02f_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_tests.log 2>&1; tail -1 gpurun_out/r02f_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r02f_smoke.log 2>&1; tail -2 gpurun_out/r02f_smoke.log
timeout 900 python bench.py > gpurun_out/r02f_bench_coarse_fine.json 2> gpurun_out/r02f_bench.err
timeout 600 python bench.py --workload fine --no-cpu-baseline > gpurun_out/r02f_bench_fine.json 2>> gpurun_out/r02f_bench.err
timeout 300 python bench.py --workload gridpool --steps 20 --warmup 3 > gpurun_out/r02f_bench_gridpool.json 2>> gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2>> gpurun_out/r02f_bench.err
for f in coarse_fine fine gridpool reference; do
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02f_bench_{f}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches_per_step"), r.get("frac"),
          (r.get("step") or {}).get("frac"), (r.get("family") or {}).get("time_weighted_frac"),
          {k: v.get("value") if isinstance(v, dict) else v for k, v in (d.get("gpu_eager_baseline") or {}).items()} if isinstance(d.get("gpu_eager_baseline"), dict) else None,
          (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), (d.get("clocks") or {}))
except Exception as e:
    print(f, "FAILED", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline --no-parity > gpurun_out/r02f_prof.log 2>&1
python profiles/summarize.py launches gpurun_out/r02f_launches.csv gpurun_out/r02f_launches.md; head -30 gpurun_out/r02f_launches.md
gzip -9 -f gpurun_out/r02f_launches.csv
