#!/usr/bin/env python
"""Static SASS inspection: opcode mix of every loop (backward branch) of one kernel in a .so.
usage: sass_loops.py <lib.so> <mangled-kernel-substring> [--dump lo hi]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ins, on = [], False
for l in out.splitlines():
    if 'Function :' in l:
        on = pat in l
        if on: ins = []; print(l.strip())
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def opc(s):
    m = re.match(r'(@!?U?P[T\d]+\s+)?([A-Z0-9_]+)', s); return m.group(2) if m else s.split()[0]
print('total', len(ins), dict(collections.Counter(opc(s) for _, s in ins).most_common(14)))
idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, s) in enumerate(ins):
    if 'BRA' in s:
        m = re.search(r'0x([0-9a-f]+)', s)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx:
            j = idx[int(m.group(1), 16)]
            c = collections.Counter(opc(x) for _, x in ins[j:i + 1])
            print('loop %x..%x len %d' % (ins[j][0], a, i - j + 1), dict(c.most_common(14)))
if '--dump' in sys.argv:
    k = sys.argv.index('--dump')
    lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
    for a, s in ins:
        if lo <= a <= hi: print('%04x %s' % (a, s))
