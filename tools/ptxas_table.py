#!/usr/bin/env python
"""python -m latticeboltzmanncuda_b200.build --force -v 2>&1 | python tools/ptxas_table.py [filter]
one line per kernel: registers, spill bytes, demangled name"""
import re
import subprocess
import sys

text = sys.stdin.read()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
cur = None
for line in text.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
    if m:
        cur = {"name": m.group(1), "regs": None, "spill": 0, "stack": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"], cur["spill"] = int(m.group(1)), int(m.group(2))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.split("\n")
for r, n in zip(rows, names):
    n = re.sub(r"lbm::\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(lbm::\w+\)", "", n)
    if flt in n:
        print("%3s regs %4d spill %4d stack  %s" % (r["regs"], r["spill"], r["stack"], n[:150]))
