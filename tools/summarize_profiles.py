"""Turns the raw captures in gpurun_out/ (ncu launch list, ncu --set full report, bench JSON) into the tracked
summaries under profiles/:  python tools/summarize_profiles.py r1"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
out = [f"# profiles {tag}: scheduler tick on 1 x B200 (cfg2-M1: 1M tasks x 256 workers x R=4, Q=16, 8 levels)", ""]

# ---- launch list (gpu__time_duration per launch, cold cache, serialised: compare SHARES) -------------
lp = os.path.join(src, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        try:
            agg.setdefault(r[ki].split("(")[0].replace("<unnamed>::", ""), []).append(float(r[vi].replace(",", "")) / 1000.0)
        except ValueError:
            pass
    tick = {k: v for k, v in agg.items() if any(x in k for x in ("tick_k", "count_k", "solve_k", "emit_k"))}
    tot = sum(sum(v) / len(v) for v in tick.values())
    out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `HQS_DEBUG_NO_COOP=1` so that",
            "ncu's kernel replay sees the otherwise cooperative tick kernel; cold cache, serialised: compare SHARES)", "",
            "| kernel | launches | avg us | share of the tick |", "|---|---|---|---|"]
    for k, v in agg.items():
        share = f"{100 * (sum(v) / len(v)) / tot:.1f} %" if k in tick else "(maintenance)"
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {share} |")
    out.append("")
    with open(os.path.join(dst, f"{tag}_launches.csv"), "w") as f:
        f.write(open(lp).read())

# ---- full-set capture ---------------------------------------------------------------------------------
rp = os.path.join(src, f"prof_{tag}.ncu-rep")
metrics = {}
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
    out += ["## `ncu --set full --clock-control none --import-source on` (one launch per kernel)", "",
            "| metric | " + " | ".join(f"`{r[hdr.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '')}`" for r in rows[2:]) + " |",
            "|---|" + "---|" * len(rows[2:])]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        def val(m):
            return float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] else None
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        unit = units[hdr.index("dram__bytes_read.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        metrics[name.split("<")[0].strip()] = {"dram_bytes_read": rd * scale if rd is not None else None,
                                       "dram_bytes_write": (wr or 0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}.get(units[hdr.index("dram__bytes_write.sum")], 1),
                                       "duration_us": val("gpu__time_duration.sum")}
    out.append("")
    json.dump(metrics, open(os.path.join(dst, f"{tag}_ncu_metrics.json"), "w"), indent=1)

bp = os.path.join(src, f"bench_{tag}.json")
if os.path.exists(bp):
    line = [l for l in open(bp).read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    json.dump(d, open(os.path.join(dst, f"{tag}_bench.json"), "w"), indent=1)
    out += ["## bench.py (not under a profiler)", "", "```", json.dumps({k: d[k] for k in ("value", "unit", "ms_per_step", "gpu_launches", "clocks")}),
            "kernels: " + json.dumps(d.get("kernels")), "roofline: " + json.dumps(d.get("roofline")), "e2e: " + json.dumps(d.get("e2e")),
            "cpu_baseline: " + json.dumps(d.get("cpu_baseline")), "extra: " + json.dumps(d.get("extra")), "```", ""]
open(os.path.join(dst, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
