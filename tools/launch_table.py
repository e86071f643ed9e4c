#!/usr/bin/env python
"""Print the last N launches of an ncu --csv metrics log: python tools/launch_table.py file.csv [N]"""
import csv, sys
from collections import OrderedDict
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
d = OrderedDict()
for r in csv.DictReader(lines):
    key = (r["ID"], r["Kernel Name"].split("(")[0][-40:])
    d.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
for (i, name), m in list(d.items())[-n:]:
    print(f"{name:42s} {m.get('gpu__time_duration.sum', 0)/1e3:8.1f} us  rd {m.get('dram__bytes_read.sum', 0)/1e6:7.1f} MB  "
          f"wr {m.get('dram__bytes_write.sum', 0)/1e6:7.1f} MB  inst {m.get('smsp__inst_executed.sum', 0)/1e6:7.2f} M")
