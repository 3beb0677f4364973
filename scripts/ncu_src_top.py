"""Summarise an `ncu --page source --csv` export: stall mix and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
totinst = sum(int(r[ix['Instructions Executed']]) for r in data)
print('total samples', tot, 'warp instructions executed', totinst)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:12]:
    print('%-28s %7d %5.1f%%' % (h, v, 100 * v / tot))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:n]
print()
for i in sorted(top):
    r = data[i]
    st = {h: int(r[ix[h]]) for h in stalls}
    best = sorted(st.items(), key=lambda x: -x[1])[:2]
    print('%4d %6s %5.2f%% exec=%9s  %-72s %s' % (i, r[ix['# Samples']], 100 * int(r[ix['# Samples']]) / tot,
          r[ix['Instructions Executed']], r[ix['Source']].strip()[:72], best))
