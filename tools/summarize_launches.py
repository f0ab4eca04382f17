"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals/shares."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        short = re.sub(r"\(.*", "", name)
        if "gemm_tc_kernel" in name:
            m = re.search(r"gemm_tc_kernel<dlc::(\w+)<([^>]*)>", name) or re.search(r"(\w+Policy)<([^>]*)>", name)
            short = "gemm_tc_kernel<%s<%s>>" % (m.group(1), m.group(2)) if m else "gemm_tc_kernel<?>"
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-58s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %5d %12.1f %10.1f %6.1f%%" % (k[:58], n, t, t / n, 100 * t / tot))
    print("total_us %.1f over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
