"""print selected raw metrics of every kernel in an .ncu-rep:  python tools/ncuq.py report.ncu-rep substr [substr ...]"""
import csv, sys, subprocess
rep = sys.argv[1]; pats = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; units = rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:80])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats): print('  %-90s %s %s' % (h, r[i], units[i]))
