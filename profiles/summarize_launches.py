"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][-48:]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        a = agg.setdefault(k, [0, 0.0, row.get("Grid Size"), row.get("Block Size")])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-50s %5s %11s %10s %7s  %s" % ("kernel", "n", "total_ms", "mean_us", "share", "grid x block"))
    for k, a in agg.items():
        print("%-50s %5d %11.3f %10.1f %6.1f%%  %s x %s" % (k, a[0], a[1] / 1e3, a[1] / a[0], a[1] / tot * 100, a[2], a[3]))


if __name__ == "__main__":
    main(sys.argv[1])
