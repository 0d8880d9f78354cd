#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
per kernel family: launches, total microseconds, share, DRAM bytes.  usage: ncu_launch_summary.py launches.csv [first] [last]"""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        k = r["ID"]
        d = per.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size"), "block": r.get("Block Size")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit == "ns" else (v if unit in ("us", "usecond") else v * 1e3 if unit == "ms" else v)
        else:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d[m] = v * scale
    return list(per.values())


def family(name):
    m = re.match(r"(?:void )?(?:b200::)?(?:\(anonymous namespace\)::)?([A-Za-z_0-9]+)(<.*)?", name)
    fam = m.group(1) if m else name
    if "Fq2T" in name:
        fam += "<Fq2>"
    elif "FqParams" in name:
        fam += "<Fq>"
    elif "FrParams" in name:
        fam += "<Fr>"
    return fam


def main():
    rows = load(sys.argv[1])
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    last = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    rows = rows[first:last]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(family(r["name"]), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += r.get("us", 0.0)
        a["rd"] += r.get("dram__bytes_read.sum", 0.0)
        a["wr"] += r.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print(f"launches {len(rows)}  total {tot:.1f} us (serialised, cold cache)")
    print(f"{'kernel':44s} {'n':>4s} {'us':>10s} {'share':>7s} {'DRAM rd MB':>11s} {'wr MB':>9s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:44s} {a['n']:4d} {a['us']:10.1f} {100 * a['us'] / tot:6.1f}% {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:9.1f}")


if __name__ == "__main__":
    main()
