#!/usr/bin/env python
"""Executed-instruction totals per kernel section (split at BAR.SYNC) from `ncu --page source --csv`."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
nel = float(sys.argv[2])
hdr = rows[1]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
sec, secs = 0, collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) < len(hdr):
        if secs: break
        continue
    src = r[iS].strip()
    m = re.match(r'(@!?U?P[T\d]+\s+)?([A-Z0-9_]+)', src)
    op = m.group(2) if m else src.split()[0]
    e, s = int(r[iE] or 0), int(r[iSamp] or 0)
    secs[sec][0] += e; secs[sec][1] += s; secs[sec][2][op] += e
    if op == 'BAR': sec += 1
tot = sum(v[0] for v in secs.values()); tots = sum(v[1] for v in secs.values())
for k, (e, s, c) in secs.items():
    print('section %d: %.1f thread-instr/element (%.1f%%), samples %.1f%%' % (k, e * 32 / nel, 100. * e / tot, 100. * s / tots))
    print('    ', ', '.join('%s %.1f' % (o, n * 32 / nel) for o, n in c.most_common(16)))
