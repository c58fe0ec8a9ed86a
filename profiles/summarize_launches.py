"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"][:90]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised)")
    print(f"# {'total ms':>10} {'launches':>8} {'us/launch':>10} {'share':>6}  kernel")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v / 1e6:12.3f} {cnt[k]:8d} {v / cnt[k] / 1e3:10.1f} {100 * v / total:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
