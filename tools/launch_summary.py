"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    n = 0
    for r in rows[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void kit::", "").replace("kit::", "")
        t = float(r[ci["Metric Value"]])
        unit = r[ci["Metric Unit"]]
        t = t / 1000 if unit == "ns" else t * 1000 if unit == "ms" else t
        agg[name][0] += 1
        agg[name][1] += t
        tot += t
        n += 1
    print(f"{n} launches, {tot:.1f} us")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:80]}` | {c} | {t:.0f} | {t / tot * 100:.1f}% | {t / c:.1f} |")


if __name__ == "__main__":
    main()
