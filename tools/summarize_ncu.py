"""Turn ncu CSV logs into the small summaries committed under profiles/.

    python tools/summarize_ncu.py launches  gpurun_out/launches_r1_small.csv  [--skip N]   -> kernel-name totals and shares
    python tools/summarize_ncu.py traffic   gpurun_out/prefix_traffic_r1.csv  --first 660 --count 660
                                                                                            -> per-launch DRAM traffic / duration
"""
import argparse
import csv
import json
import re
import sys


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    return rows


def num(s):
    return float(s.replace(",", ""))


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.replace("void ", "").strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "traffic"])
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=0)
    a = ap.parse_args()
    rows = read(a.csv)
    if a.mode == "launches":
        # one row per (launch, metric); gpu__time_duration.sum in ns (or us, see unit column)
        per = {}
        order = []
        for r in rows:
            if r["Metric Name"] != "gpu__time_duration.sum":
                continue
            lid = int(r["ID"])
            unit = r["Metric Unit"]
            v = num(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
            order.append((lid, short(r["Kernel Name"]), v))
        order = [o for o in order if o[0] >= a.skip]
        tot = sum(v for _, _, v in order)
        for _, n, v in order:
            d = per.setdefault(n, [0, 0.0])
            d[0] += 1
            d[1] += v
        out = {"launches": len(order), "total_us": tot,
               "kernels": [{"kernel": n, "launches": c, "total_us": round(v, 1), "share": round(v / tot, 4)}
                           for n, (c, v) in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]]}
        json.dump(out, sys.stdout, indent=1)
    else:
        per = {}
        for r in rows:
            lid = int(r["ID"])
            unit = r["Metric Unit"]
            v = num(r["Metric Value"])
            if r["Metric Name"] == "gpu__time_duration.sum":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
            else:
                v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per.setdefault(lid, {})[r["Metric Name"]] = v
        ids = sorted(per)
        if a.count:
            ids = ids[a.first:a.first + a.count]
        rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
        wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)
        us = sum(per[i]["gpu__time_duration.sum"] for i in ids)
        json.dump({"launches": len(ids), "first_launch": ids[0], "last_launch": ids[-1],
                   "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
                   "dram_bytes_per_launch": (rd + wr) / len(ids), "duration_us_total_under_ncu": us,
                   "duration_us_per_launch_under_ncu": us / len(ids)}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
