#!/usr/bin/env python3
"""Turn gpurun_out/ ncu outputs into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <launches.csv> <prof.ncu-rep> <tag>"""
import csv
import collections
import subprocess
import sys

launches, rep, tag = sys.argv[1:4]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
ui = H.index("Metric Unit")
tot = collections.OrderedDict()
cnt = collections.Counter()
allv = {}
for r in rows[hdr + 1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns":
        v /= 1e3
    elif r[ui] == "ms":
        v *= 1e3
    name = r[ki].split("(")[0].replace("void b200::", "").replace("b200::", "")
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
    allv.setdefault(name, []).append(v)
total = sum(tot.values())
out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)", "",
       f"source: `{launches}`; command: see the session script named in the tag's bench notes (`ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py ...`)",
       "Per-launch times under ncu are cold-cache and serialised: read the SHARE column, not the absolute.",
       "`working` = launches that did their work (longer than a third of the kernel's longest launch): a graph replay carries spare",
       "passes and speculative checks whose kernels return at once when they are not due, which drags the plain average down.", "",
       "| kernel | launches | total us | avg us | working launches | avg us of those | share |", "|---|---:|---:|---:|---:|---:|---:|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    w = [x for x in allv[k] if x > max(allv[k]) / 3.0]
    out.append(f"| `{k}` | {cnt[k]} | {v:.1f} | {v / cnt[k]:.2f} | {len(w)} | {sum(w) / len(w):.2f} | {100 * v / total:.1f}% |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
seen = set()
out = [f"# {tag}: ncu --set full, one launch per kernel", "",
       f"source: `{rep}` (not committed, 10 MB); `ncu --set full --clock-control none --import-source on -k regex:spmv_sell|primal_step|step_rule`",
       "workload: S3 (m=n=1e6, nnz=8e6, fp64), launches taken inside the PDHG pass sequence.", ""]
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    full = r[idx["Kernel Name"]][:70]
    if full in seen:
        continue
    seen.add(full)
    out += [f"## `{full}`", "", "| metric | value |", "|---|---|"]
    for w in want:
        if w in idx:
            out.append(f"| {w} | {r[idx[w]]} {units[idx[w]]} |")
    try:
        rd = float(r[idx['dram__bytes_read.sum']].replace(',', '')); wr = float(r[idx['dram__bytes_write.sum']].replace(',', ''))
        u = units[idx['dram__bytes_read.sum']]
        out.append(f"| **traffic = dram read + write** | {rd + wr:.1f} {u} |")
    except Exception:
        pass
    out.append("")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_launches.md", f"profiles/{tag}_ncu_summary.md")
