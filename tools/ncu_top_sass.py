"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source sass ...` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(x):
    try: return float(x.replace(',', ''))
    except Exception: return 0.0
S = idx['# Samples']
tot = sum(f(r[S]) for r in data)
print('kernel', rows[0][1][:80], 'total samples', tot, 'instructions', len(data))
order = sorted(range(len(data)), key=lambda i: -f(data[i][S]))[:n]
for i in sorted(order):
    r = data[i]
    print('%6.2f%%  #%5d  %s' % (100 * f(r[S]) / tot, i, r[idx['Source']][:100]))
