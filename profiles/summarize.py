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


def mix(src, dst):
    """Append the dynamic instruction mix, the issue utilisation and the stall reasons of every kernel of an ncu report
    (raw + source pages): where the issue slots of a kernel go."""
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr = rows[0]
    out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(out)))
    with open(dst, "a") as f:
        for r in rows[2:]:
            f.write(f"### `{short(r[hdr.index('Kernel Name')])}`: issue slots and stalls\n\n")
            for m in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum"):
                if m in hdr:
                    f.write(f"* {m} = {r[hdr.index(m)]}\n")
            st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i] or 0))
                  for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
            st.sort(key=lambda kv: -kv[1])
            f.write("* stalls per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in st[:8]) + "\n")
        # instruction mix of the first kernel of the source page
        h2 = None
        agg, tot = collections.Counter(), 0
        for r in srows:
            if h2 is None:
                if "Instructions Executed" in r:
                    h2 = r
                    ia, isrc = r.index("Instructions Executed"), r.index("Source")
                continue
            if len(r) <= ia or not r[ia].isdigit():
                continue
            src_txt = r[isrc].strip().split()
            if not src_txt:
                continue
            op = (src_txt[1] if src_txt[0].startswith("@") and len(src_txt) > 1 else src_txt[0]).split(".")[0]
            agg[op] += int(r[ia])
            tot += int(r[ia])
        if tot:
            f.write("* dynamic instruction mix (warp instructions, all kernels of the report): " +
                    ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in agg.most_common(16)) + f"; total {tot}\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "mix": mix}[sys.argv[1]](sys.argv[2], sys.argv[3])
