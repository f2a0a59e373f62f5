"""Per-kernel table from one ncu launch list (csv) and, optionally, the key columns of an `ncu --set full` report.

    python profiles/summarize_list.py <launches.csv> <out.md> "<title>" [<report.ncu-rep> <out.csv>]
"""
import collections, csv, subprocess, sys
src, out, title = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name, val, unit = r[4].split("(")[0], float(r[-1]), r[-2]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1000.0
    per.setdefault(name, []).append(us)
tot = sum(sum(v) for v in per.values())
with open(out, "w") as f:
    f.write(f"# {title}\n\nPer-launch times are cold-cache and serialised: compare SHARES, not absolutes. Total {tot / 1000:.2f} ms.\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| {k[:110]} | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f}% |\n")
if len(sys.argv) > 5:
    rep, outcsv = sys.argv[4:6]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
    hdr, units = rr[0], rr[1]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
            "launch__cluster_dim_x", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]
    idx = [i for i, h in enumerate(hdr) if h in keep]
    with open(outcsv, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rr[2:]:
            w.writerow([r[i] for i in idx])
