#!/usr/bin/env python
"""Summarise ncu artefacts into profiles/: (1) launch list -> per-kernel time shares, (2) --set full raw page -> key counters.
usage: summarize_profiles.py <tag> <launches.csv> <full.ncu-rep>"""
import collections, csv, os, re, subprocess, sys
tag, launches, rep = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    ns = v * 1e3 if u.startswith("us") else (v * 1e6 if u.startswith("ms") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("tpat::", "").replace("void ", "")
    agg[name][0] += 1; agg[name][1] += ns; tot += ns
with open(os.path.join(out, f"{tag}_launch_shares.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 3 --warmup 3 (launches 300..399 = one forward + 9)\n")
    f.write(f"# cold-cache, serialised: compare SHARES, not absolutes.  total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}%  n={n:3d}  {k}\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(os.path.join(out, f"{tag}_ncu_full_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on; tools/prof_kernels.py at B=64, N=513 (headline shapes)\n")
    for d in data:
        f.write(f"--- {d[idx['Kernel Name']][:110]}\n")
        for k in keys:
            if k in idx: f.write(f"    {k:72s} {d[idx[k]]:>18s} {units[idx[k]]}\n")
print(open(os.path.join(out, f"{tag}_launch_shares.txt")).read())
# dram traffic per launch of the kernels bench.py reports a roofline for -> profiles/ncu_traffic.json (read by bench.py)
import json
names = {"gemm_tc2_kernel<1": "gemm_fc1_tcgen05", "layernorm_kernel": "layernorm", "attention_tc_kernel<0>": "attention_tcgen05"}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1))
traffic = {}
for d in data:
    kn = d[idx["Kernel Name"]]
    for pat, key in names.items():
        if pat in kn and key not in traffic:
            rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            traffic[key] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "kernel": kn[:100], "source": f"profiles/{tag}_ncu_full_summary.txt"}
json.dump(traffic, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
