"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | min us | max us | share of kernel time |")
    print("|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {sum(v) / tot:.3f} |")
    print(f"\ntotal kernel time {tot:.1f} us over {sum(len(v) for v in agg.values())} launches "
          "(ncu: serialised, cold caches — compare shares, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
