"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def summarize(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        v = float(r["Metric Value"])
        v = v / 1000 if r["Metric Unit"] == "ns" else v * 1000 if r["Metric Unit"] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"launches {sum(v[0] for v in agg.values())}  total {tot:.1f} us (cold-cache, serialised under ncu: compare SHARES)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]:10.1f} us {v[1] / tot * 100:5.1f}%  n={v[0]:5d}  avg {v[1] / v[0]:7.2f} us  {k[:100]}")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
