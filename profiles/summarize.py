"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/r01_launches_X.md
    python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/r01_full_X.md
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "launch__occupancy_limit_registers",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:90]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", "")) / 1e3          # ns -> us
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} us over {sum(n for n, _ in agg.values())} launches\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[hdr.index('Kernel Name')])}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in KEEP:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
