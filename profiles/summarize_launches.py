"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols, data = rows[hdr], rows[hdr + 1:]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised)")
    print("# share   launches   avg_us   kernel")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / tot * 100:6.2f}%  {n:6d}  {v / n / 1e3:9.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
