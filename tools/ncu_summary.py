"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            unit = d["Metric Unit"]
            v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
            agg.setdefault(d["Kernel Name"].split("(")[0][:60], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-62s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-62s %5d %12.1f %10.1f %6.1f%%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("%-62s %5s %12.1f" % ("TOTAL", "", tot))


if __name__ == "__main__":
    main(sys.argv[1])
