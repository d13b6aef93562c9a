"""Top SASS lines by warp-stall samples from `ncu --page source --csv` output: python tools/ncu_hot.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
tot = 0
for idx, r in enumerate(rows[2:]):
    try:
        v = int(r[ci['# Samples']])
    except Exception:
        continue
    tot += v
    data.append((v, idx, r))
print('total samples', tot, 'instructions', len(data))
for v, idx, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((int(r[ci[s]] or 0), s) for s in st), reverse=True)[:2]
    print("%6d  #%5d  exec %8s  %-60s %s" % (v, idx, r[ci['Instructions Executed']], r[ci['Source']].strip()[:60], ' '.join('%s=%d' % (s[6:], c) for c, s in top if c)))
