"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python profiles/summarize.py <tag>        # e.g. r1  -> reads gpurun_out/launches_<tag>.csv, prof_umma_<tag>.ncu-rep
"""
import csv, collections, os, subprocess, sys, json
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name, val, unit = r[4].split("(")[0], float(r[-1]), r[-2]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1000.0
    per.setdefault(name, []).append(us)
tot = sum(sum(v) for v in per.values())
with open(os.path.join(root, "profiles", f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 1`\n\n")
    f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| {k} | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {100*sum(v)/tot:.1f}% |\n")
rep = os.path.join(root, "gpurun_out", f"prof_umma_{tag}.ncu-rep")
rawcsv = os.path.join(root, "gpurun_out", f"prof_umma_{tag}_raw.csv")   # `ncu -i ... --page raw --csv` run on the GPU box
                                                                         # (a --set full report of a whole step exceeds gpurun's 64 MiB)
if os.path.exists(rep) or os.path.exists(rawcsv):
    raw = (subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
           if os.path.exists(rep) else open(rawcsv).read())
    rr = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
    hdr, units = rr[0], rr[1]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
            "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]
    idx = [i for i, h in enumerate(hdr) if h in keep]
    with open(os.path.join(root, "profiles", f"{tag}_umma_ncu_full.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rr[2:]:
            w.writerow([r[i] for i in idx])
    # per-launch DRAM traffic of the dominant kernel for bench.py's roofline.traffic
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    def to_bytes(v, u):
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(v) * m.get(u, 1)
    out = {}
    for r in rr[2:]:
        nm = r[ik].split("(")[0].replace("void ", "").replace("drb::", "")
        key = ("umma_gate_dual_kernel" if "gate_pers_kernel<3, 1>" in nm or "gate_pers_kernel<1, 1>" in nm or "gate_n4_kernel<1>" in nm
               else "umma_gate_n4_kernel" if "gate_n4" in nm else "umma_gate_kernel" if "gate" in nm
               else "umma_res_kernel" if "res_pers" in nm else "umma_zgemm_kernel" if "zgemm" in nm else nm)
        out.setdefault(key + "_dram_bytes_per_launch", []).append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
    out = {k: sum(v) / len(v) for k, v in out.items()}
    out["source"] = f"ncu --set full --clock-control none, profiles/{tag}_umma_ncu_full.csv"
    json.dump(out, open(os.path.join(root, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(open(os.path.join(root, "profiles", f"{tag}_launches_summary.md")).read())
