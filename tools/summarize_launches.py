"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        tot[name][0] += 1
        tot[name][1] += ns
    total = sum(v[1] for v in tot.values())
    lines = [f"total kernel time {total/1e6:.2f} ms over {sum(v[0] for v in tot.values())} launches (cold-cache, serialised: compare shares)"]
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{ns/total*100:6.2f}%  {ns/1e6:9.3f} ms  {n:6d} launches  {ns/n/1e3:9.2f} us/launch  {name}")
    text = "\n".join(lines)
    print(text)
    if out:
        with open(out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
