"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = d["Kernel Name"][:90]
            v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| launches | total us | avg us | share | kernel |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f} % | `{k}` |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
