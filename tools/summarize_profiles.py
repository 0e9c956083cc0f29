#!/usr/bin/env python
"""Turn ncu captures into the tracked summaries under profiles/.

    python tools/summarize_profiles.py full  gpurun_out/r02_kernels.ncu-rep   profiles/r02_ncu_full_summary.txt
    python tools/summarize_profiles.py fwd   gpurun_out/launches_fwd_dram.csv profiles/ncu_traffic.json

`full`: per kernel of an `ncu --set full` report: duration, tensor / XU pipe activity, issue activity, achieved occupancy,
        registers, DRAM bytes and throughput, L2 throughput.
`fwd` : from a launch list with dram__bytes_* metrics over `bench.py --no-graph`: DRAM bytes of ONE forward (the last
        complete one) in total and for the kernels bench.py names in its roofline entries -> profiles/ncu_traffic.json."""
import csv
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
]


def full(rep, out):
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(m for m, _ in METRICS)]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"source: {rep} (ncu --set full --clock-control none, one launch per kernel at B = 64, N = 513 unless noted)",
             "units: " + ", ".join(f"{lab} [{units[ix[m]]}]" for m, lab in METRICS if m in ix), ""]
    lines.append("kernel".ljust(58) + "".join(lab.rjust(11) for _, lab in METRICS))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void ", "")
        name = name[:name.find("(")] if "(" in name else name
        vals = []
        for m, _ in METRICS:
            v = r[ix[m]] if m in ix else ""
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
            vals.append(v.rjust(11))
        lines.append(name[:57].ljust(58) + "".join(vals))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def fwd(csv_path, out):
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = {}
    order = []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in per:
            per[i] = {"name": r["Kernel Name"]}
            order.append(i)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1)
        per[i][r["Metric Name"]] = v * scale
    starts = [k for k, i in enumerate(order) if "patchify_kernel" in per[i]["name"]]
    a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(order))
    step = [per[i] for i in order[a:b]]
    tot = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in step)

    def first(pred):
        for k in step:
            if pred(k["name"]):
                return int(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0))
        return None
    res = {
        "_source": f"{csv_path}: launches {a}..{b} = one forward of bench.py --no-graph (AudioMAE 1024x128, B = 64, keep 0.7); "
                   "dram__bytes_read.sum + dram__bytes_write.sum per launch",
        "forward_total": {"dram_bytes": int(tot), "launches": len(step)},
        "gemm_fc1_tcgen05": {"dram_bytes": first(lambda n: "gemm_tc2_kernel<1" in n)},
        "gemm_fc2_tcgen05": {"dram_bytes": first(lambda n: "gemm_tc2_kernel<2, float, 0" in n)},
        "layernorm": {"dram_bytes": first(lambda n: "layernorm_kernel" in n and "gather" not in n)},
        "attention_tcgen05": {"dram_bytes": first(lambda n: "attention_tc4_kernel" in n or "attention_tc_kernel<0" in n)},
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"full": full, "fwd": fwd}[sys.argv[1]](sys.argv[2], sys.argv[3])
