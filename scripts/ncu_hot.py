"""Hottest SASS lines of an `ncu --page source --csv` export: ncu_hot.py file.csv [kernel_index] [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
s = starts[k]
e = starts[k + 1] if k + 1 < len(starts) else len(rows)
print(rows[s][1])
hdr = rows[s + 1]
body = [r for r in rows[s + 2:e] if len(r) >= 6]
ci = hdr.index('Warp Stall Sampling (All Samples)')
total = sum(int(r[ci] or 0) for r in body)
print('total samples', total)
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][ci] or 0))[:top]
for i, r in sorted(ranked):
    print(f'{i:5d} {int(r[ci] or 0):7d} {100.0 * int(r[ci] or 0) / max(total, 1):5.1f}%  {r[1][:110]}')
