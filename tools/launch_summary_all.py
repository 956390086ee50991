"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel over the whole capture
(tools/launch_summary.py windows one SNDCGAN step; this one serves captures of any workload).
usage: python tools/launch_summary_all.py launches.csv [steps]"""
import collections
import csv
import sys

from launch_summary import short


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for row in r:
        try:
            k, v = short(row[ki]), float(row[vi].replace(",", ""))
        except Exception:
            continue
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print("%d launches, %.3f ms device time in the capture (%.3f ms / step over %d steps; ncu times are cold-cache, "
          "serialised)" % (n, total / 1e6, total / 1e6 / steps, steps))
    print("| kernel | launches/step | us/step | share |")
    print("|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %.1f | %.1f | %.1f%% |" % (k, c / steps, t / 1e3 / steps, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
