#!/usr/bin/env python
"""sass_hot.py FILE.src.csv[.gz] -- digest of an `ncu --page source --csv --print-source sass` dump:
stall-reason totals, executed-instruction mix, and the hottest contiguous SASS regions."""
import csv, gzip, sys, collections, re
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
f = gzip.open(path, 'rt') if path.endswith('.gz') else open(path)
rows = list(csv.reader(f))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = rows[2:]
tot = collections.Counter(); mix = collections.Counter()
samples = 0; inst = 0
for r in body:
    s = int(r[ix['# Samples']]); e = int(r[ix['Instructions Executed']])
    samples += s; inst += e
    for c in stall_cols:
        tot[c] += int(r[ix[c]])
    if e:
        op = r[1].split()
        op = [o for o in op if not o.startswith('@')][0].split('.')[0]
        mix[op] += e
print("static instructions %d, executed warp-instructions %d, samples %d" % (len(body), inst, samples))
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(samples, 1)) for k, v in tot.most_common(8)))
print("mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / inst) for k, v in mix.most_common(16)))
# hot regions: windows of executed instructions
exec_rows = [(i, r) for i, r in enumerate(body) if int(r[ix['Instructions Executed']]) > 0]
print("executed static instructions: %d" % len(exec_rows))
W = 64
best = []
for k in range(0, len(exec_rows), W):
    chunk = exec_rows[k:k + W]
    s = sum(int(r[ix['# Samples']]) for _, r in chunk)
    best.append((s, k))
best.sort(reverse=True)
for s, k in best[:top]:
    chunk = exec_rows[k:k + W]
    c = collections.Counter()
    for _, r in chunk:
        for col in stall_cols:
            c[col] += int(r[ix[col]])
    e = max(int(r[ix['Instructions Executed']]) for _, r in chunk)
    ops = collections.Counter(r[1].split()[0].split('.')[0] if not r[1].split()[0].startswith('@') else r[1].split()[1].split('.')[0] for _, r in chunk)
    print("  rows %6d-%6d: %5.1f%% of samples, max exec %d; %s | %s" % (chunk[0][0], chunk[-1][0], 100.0 * s / samples, e,
          ", ".join("%s %d" % (a[6:], b) for a, b in c.most_common(4)), ", ".join("%s %d" % ab for ab in ops.most_common(6))))
