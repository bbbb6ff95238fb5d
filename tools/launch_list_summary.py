"""Per-kernel totals and shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`), for comparison with the
CUDA-event shares bench.py prints (per-launch times under ncu are serialised and cold-cache: the SHARES must agree).
    python tools/launch_list_summary.py profiles/r02_launches.csv [steps] > profiles/r02_launches_summary.txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows:
    name = re.sub(r"<.*", "", r[4].split("(")[0]).replace("void ", "").replace("kdcc::", "").strip()
    try:
        v = float(r[-1].replace(",", ""))
    except ValueError:
        continue
    unit = r[-2]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1e-3)
    tot[name] = tot.get(name, 0.0) + us
    cnt[name] += 1
total = sum(tot.values())
print("# source: %s  (%d launches, %d timed steps; serialised ncu durations)" % (path, len(rows), steps))
print("%-44s %9s %12s %8s %12s" % ("kernel", "launches", "us total", "share", "us / step"))
for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-44s %9d %12.1f %8.3f %12.1f" % (name[:44], cnt[name], us, us / total, us / steps))
print("%-44s %9d %12.1f %8.3f %12.1f" % ("total", len(rows), total, 1.0, total / steps))
