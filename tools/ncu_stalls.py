"""Summarise the source page of an ncu report: top SASS instructions by stall samples, with the reasons.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [top_n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    samp = ci["# Samples"]
    total = sum(int(r[samp] or 0) for r in body)
    print("total samples", total, "instructions", len(body))
    agg = {}
    for r in body:
        for h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[ci[h]] or 0)
    print("by reason:", ", ".join(f"{k[6:]}={v * 100 // max(total, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 >= total))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        n = int(r[samp] or 0)
        reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {n * 100.0 / max(total, 1):5.1f}%  {r[ci['Source']][:90]:90s} " + " ".join(f"{k}={v}" for v, k in reasons if v))


if __name__ == "__main__":
    main()
