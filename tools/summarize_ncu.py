#!/usr/bin/env python
"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_ncu.py r01 gpurun_out/launches_r01.csv gpurun_out/prof_sweep_r01.ncu-rep c5g7-2d
"""
import csv, io, json, os, re, subprocess, sys
from collections import OrderedDict

tag, launches_csv, rep, workload = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---------------------------------------------------------------- launch list
rows = [r for r in csv.reader(l for l in open(launches_csv) if not l.startswith("==")) if r]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1; c[1] += us
total = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: launch list of `bench.py --steps 5 --warmup 3` ({workload}), ncu `gpu__time_duration.sum`, "
            "`--clock-control none`\n\nPer-launch times are cold-cache and serialised: compare SHARES.\n\n"
            "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.2f} % |\n")
    f.write(f"\nTotal {total / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches.\n")

# ---------------------------------------------------------------- full capture
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u, v = rr[0], rr[1], rr[2]
m = {a: (c, b) for a, b, c in zip(h, u, v)}
def g(name):
    return m.get(name, ("n/a", ""))
keys = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"), ("launch__grid_size", "grid size"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (none expected)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed_op_global_red.sum", "RED (tally) warp instructions"),
    ("lts__t_sectors_op_red.sum", "L2 sectors, reductions"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]
name = g("Kernel Name")[0] if "Kernel Name" in m else "sweep kernel"
with open(os.path.join(out_dir, f"{tag}_sweep_kernel.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of the sweep kernel, workload {workload}\n\n"
            f"Kernel: `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k, label in keys:
        val, unit = g(k)
        f.write(f"| {label} (`{k}`) | {val} | {unit} |\n")
def num(name):
    val, unit = g(name)
    x = float(val.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return x * mult
traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
tj = os.path.join(out_dir, "roofline_traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d[workload] = {"dram_bytes_per_launch": traffic, "from": f"profiles/{tag}_sweep_kernel.md"}
json.dump(d, open(tj, "w"), indent=1)
print("DRAM traffic per launch: %.3f GB" % (traffic / 1e9))
