"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps_in_capture] > profiles/<round>_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        rows.append((row["Kernel Name"], row["Grid Size"].replace(" ", ""), row["Block Size"].replace(" ", ""), v))
    return rows


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0][:70]


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = load(path)
    agg = collections.OrderedDict()
    for k, g, b, v in rows:
        a = agg.setdefault(short(k), [0, 0.0, 0.0, g, b])
        a[0] += 1
        a[1] += v
        if v > a[2]:
            a[2], a[3], a[4] = v, g, b
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not k.startswith("at") and "cub" not in k.lower() and "at_cuda" not in k)
    print("# ncu launch list summary: %s" % path)
    print()
    print("%d launches, %.1f us summed device time over %d scene passes (%.1f us per pass); kernels of libseggroup_b200.so: %.1f%% of the time."
          % (len(rows), tot, steps, tot / steps, 100 * ours / tot))
    print()
    print("| kernel | launches | total us | share | slowest launch us | its grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.002:
            continue
        print("| `%s` | %d | %.1f | %.1f%% | %.1f | %s | %s |" % (k, a[0], a[1], 100 * a[1] / tot, a[2], a[3], a[4]))


if __name__ == "__main__":
    main()
