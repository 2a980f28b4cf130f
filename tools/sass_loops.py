#!/usr/bin/env python
"""Instruction mix of the innermost loops of a cuobjdump -sass listing (loops holding VIADDMNMX).
usage: sass_loops.py file.sass [min_viaddmnmx]"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
minv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ins = []
for l in lines:
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in idx: loops.append((idx[tgt], i))
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
ALU = ('VIADDMNMX', 'VIMNMX', 'LOP3', 'ISETP', 'SEL', 'SHF', 'VIADD', 'PRMT', 'IADD3', 'LEA', 'MOV', 'PLOP3', 'SGXT', 'BMSK', 'POPC', 'FLO', 'R2P', 'P2R')
for lo, hi in inner:
    body = ins[lo:hi + 1]
    nv = sum('VIADDMNMX' in t or 'VIMNMX' in t for _, t in body)
    if nv < minv: continue
    c = collections.Counter()
    for _, t in body:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        c[t.split()[0].split('.')[0] if not t.startswith('IMAD') else t.split()[0]] += 1
    alu = sum(v for k, v in c.items() if k in ALU)
    print('loop %#x..%#x  %d instr, %d minmax, ALU-pipe %d' % (ins[lo][0], ins[hi][0], len(body), nv, alu))
    print('   ' + '  '.join('%s %d' % kv for kv in c.most_common()))
