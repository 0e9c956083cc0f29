#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of ONE step.

    python tools/launch_shares.py gpurun_out/launches.csv [--last N | --from A --to B] [--group]

Without a range it finds the last complete step by looking for the first kernel of a step (patchify_kernel) and
takes the launches from the last-but-one occurrence to the last one.  Times under ncu are cold-cache and serialised:
only the SHARES are meaningful."""
import argparse
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("tpat::", "").replace("at::native::", "aten::")
    return name[:110]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--first", default="patchify_kernel", help="kernel that starts a step")
    ap.add_argument("--from", dest="a", type=int)
    ap.add_argument("--to", dest="b", type=int)
    args = ap.parse_args()
    rows = []
    with open(args.csv) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], v))
    if args.a is None:
        starts = [i for i, (_, n, _) in enumerate(rows) if args.first in n]
        if len(starts) < 2:
            a, b = 0, len(rows)
        else:
            a, b = starts[-2], starts[-1]
    else:
        a, b = args.a, args.b
    step = rows[a:b]
    total = sum(v for _, _, v in step)
    agg = OrderedDict()
    for _, n, v in step:
        k = short(n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    print(f"launches {a}..{b} of {len(rows)}: {len(step)} launches, {total / 1e3:.3f} ms (serialised, cold cache, under ncu)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.2f} %  {t:10.1f} us  x{c:<4d} {k}")


if __name__ == "__main__":
    main()
