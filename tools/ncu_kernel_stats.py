"""Per-kernel averages from an `ncu --csv --page raw` dump: duration, DRAM bytes, tensor-pipe activity.
   python tools/ncu_kernel_stats.py raw.csv [out.json]"""
import csv
import json
import re
import sys
from collections import defaultdict

WANT = {"gpu__time_duration.sum": "ns", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct"}
UNIT = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out=None):
    rows = [r for r in csv.reader(l for l in open(path, newline="") if not l.startswith("=="))]
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: defaultdict(float))
    cnt = defaultdict(int)
    for r in data:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*$", "", r[name_i]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        cnt[name] += 1
        for m, key in WANT.items():
            if m in cols and r[cols[m]] not in ("", "n/a"):
                v = float(r[cols[m]].replace(",", "")) * UNIT.get(units[cols[m]], 1)
                agg[name][key] += v
    res = {}
    for name, n in sorted(cnt.items(), key=lambda kv: -agg[kv[0]]["ns"]):
        a = agg[name]
        res[name] = {"launches": n, "avg_us": a["ns"] / n / 1e3, "avg_dram_bytes": (a["dram_read"] + a["dram_write"]) / n,
                     "tensor_pct_active": a["tensor_pct_active"] / n, "tensor_pct_elapsed": a["tensor_pct_elapsed"] / n,
                     "dram_pct": a["dram_pct"] / n, "sm_pct": a["sm_pct"] / n}
        print(f"{name[:60]:60s} n={n:4d} avg {res[name]['avg_us']:8.1f} us  dram {res[name]['avg_dram_bytes']/1e6:8.1f} MB  "
              f"tensor {res[name]['tensor_pct_active']:5.1f}% of active ({res[name]['tensor_pct_elapsed']:5.1f}% of elapsed)  "
              f"dram {res[name]['dram_pct']:5.1f}%")
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
