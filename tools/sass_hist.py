#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) and stall samples from `ncu --page source --csv`
for the first kernel in the file.  usage: sass_hist.py source.csv [n_elements]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
nel = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iT = hdr.index('Thread Instructions Executed')
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Address':
        if r and r[0] == 'Kernel Name': break
        continue
    src = r[iS].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else src
    base = op.split('.')[0]
    if base in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'MUFU'): key = base
    elif base in ('LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'ATOMS', 'ATOMG', 'RED'): key = op if base in ('LDS','STS') else base
    else: key = base
    e, s = int(r[iE] or 0), int(r[iSamp] or 0)
    ops[key] += e; samp[key] += s; tot += e; tots += s
    body.append((e, s, src))
print('total warp instr %d, samples %d' % (tot, tots))
for k, v in ops.most_common(40):
    line = '%-14s %12d %5.1f%%  samples %5.1f%%' % (k, v, 100.0 * v / tot, 100.0 * samp[k] / max(tots, 1))
    if nel: line += '   %.1f thread-instr/element' % (v * 32 / nel)
    print(line)
if '--top' in sys.argv:
    print('--- top stall lines')
    for e, s, src in sorted(body, key=lambda t: -t[1])[:40]:
        print('%8d %6d  %s' % (e, s, src))
