"""Scratch: read an `ncu --page source --csv` export: stall reasons overall and per code region (by execution count)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot)
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v: print("%-24s %7d %5.1f%%" % (s, v, 100 * v / tot))
ex = [int(r[ix['Instructions Executed']]) for r in data]
hist = collections.Counter()
for r, e in zip(data, ex):
    hist[e] += int(r[ix['# Samples']])
print("samples by execution count (top):")
for e, n in sorted(hist.items(), key=lambda kv: -kv[1])[:12]:
    print("  exec %10d  instrs %5d  samples %7d %5.1f%%" % (e, sum(1 for x in ex if x == e), n, 100 * n / tot))
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    for i in range(lo, hi):
        r = data[i]
        top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print("%5d %-62s ex=%9s smp=%6s  %s" % (i, r[ix['Source']].strip()[:62], r[ix['Instructions Executed']], r[ix['# Samples']],
              " ".join("%s=%d" % (s[6:], v) for v, s in top if v)))
