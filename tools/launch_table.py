"""one line per kernel launch of an `ncu --csv --metrics ...` log: python tools/launch_table.py launches.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
cur = {}
for r in rows[1:]:
    cur.setdefault((r[idi], r[ki][:64]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), m in cur.items():
    t = m.get("gpu__time_duration.sum", 0.0) / 1e6
    b = (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / 1e9
    print("%3s %-64s %8.3f ms %8.3f GB %7.0f GB/s" % (i, k, t, b, b / t * 1e3 if t else 0.0))
