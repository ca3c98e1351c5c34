"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"<unnamed>::|void |at::native::|\(anonymous namespace\)::", "", name)
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"total {total/1e3:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v/1e3:9.3f} ms {100*v/total:5.1f}%  n={cnt[k]:4d}  avg {v/cnt[k]:8.1f} us  {k[:110]}")
