"""Tabulate selected metrics of `ncu --page raw --csv` exports side by side: ncu_table.py a.raw.csv b.raw.csv ..."""
import csv, re, sys
def load(fn):
    rows = list(csv.reader(open(fn)))
    return {h: v for h, v in zip(rows[0], rows[2])}
ds = [load(f) for f in sys.argv[1:]]
pat = re.compile(r'(tensor|tmem|issue_stalled.*_per_issue_active|sm__cycles_elapsed.max|gpu__time|dram__bytes_(read|write).sum$|'
                 r'lts__t_bytes.sum$|lts__t_sectors_op_(read|write).sum$|throughput.avg.pct|l1tex__m_xbar2l1tex|'
                 r'l1tex__m_l1tex2xbar|shared|sm__inst_executed.sum$|hit_rate|registers|grid_size|dynamic)')
for k in ds[0]:
    if pat.search(k) and not k.startswith('device') and '.min.' not in k and '.max.pct' not in k and \
            '.sum.pct' not in k and 'per_second' not in k and 'peak_sustained_active' not in k:
        vals = [d.get(k, '') for d in ds]
        if all(v in ('', '0') for v in vals):
            continue
        def f(v):
            try:
                return f'{float(v):,.1f}'
            except ValueError:
                return v
        print(f'{k[:100]:100s} ' + ' | '.join(f(v) for v in vals))
