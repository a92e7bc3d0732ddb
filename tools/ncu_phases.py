"""Aggregate an `ncu --page source --csv` export between barriers: instructions and stall samples per phase."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
src = ci['Source']; ex = ci['Instructions Executed']; smp = ci['Warp Stall Sampling (All Samples)']
tot = sum(int(r[ex] or 0) for r in data); tots = sum(int(r[smp] or 0) for r in data)
print("total inst", tot, "samples", tots, "rows", len(data))
ph = 0; agg = collections.OrderedDict(); ops = collections.Counter()
for idx, r in enumerate(data):
    s = r[src]; n = int(r[ex] or 0); sm = int(r[smp] or 0)
    a = agg.setdefault(ph, [0, 0, 0, None, idx]); a[0] += n; a[1] += sm; a[2] += 1
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    ops[m.group(2).split('.')[0] if m else '?'] += n
    if 'BAR.SYNC' in s or 'BAR.ARV' in s or 'BAR.RED' in s:
        a[3] = s.strip(); ph += 1
for k, a in agg.items():
    print("%2d inst %5.1f%% samp %5.1f%% rows %4d..%4d end=%s" % (k, 100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[4], a[4] + a[2] - 1, a[3]))
print(ops.most_common(16))
if len(sys.argv) > 2:
    lo, hi2 = int(sys.argv[2]), int(sys.argv[3])
    for r in data[lo:hi2 + 1]:
        print(r[ex].rjust(10), r[smp].rjust(6), r[src][:110])
